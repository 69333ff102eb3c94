"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_*]) per kernel name."""
import collections
import csv
import io
import sys


def main(path, min_us=0.0):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
    rows = collections.OrderedDict()
    for r in rd:
        k = (r["ID"], r["Kernel Name"][:70])
        rows.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    tob = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tous = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}
    agg = collections.OrderedDict()
    for (_, name), m in rows.items():
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        t = m["gpu__time_duration.sum"]
        a[1] += t[0] * tous[t[1]]
        for key, slot in (("dram__bytes_read.sum", 2), ("dram__bytes_write.sum", 3)):
            if key in m:
                a[slot] += m[key][0] * tob[m[key][1]]
    total = sum(a[1] for a in agg.values())
    print("%-70s %5s %10s %10s %6s %9s %9s %8s" % ("kernel", "n", "total us", "avg us", "share", "rd MB", "wr MB", "GB/s"))
    for name, a in agg.items():
        if a[1] < min_us:
            continue
        print("%-70s %5d %10.1f %10.1f %5.1f%% %9.1f %9.1f %8.0f" % (name, a[0], a[1], a[1] / a[0], 100 * a[1] / total,
              a[2] / 1e6 / a[0], a[3] / 1e6 / a[0], (a[2] + a[3]) / a[1] / 1e3 if a[1] else 0))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.0)
