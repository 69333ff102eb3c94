"""Builds libdistgcn_b200.so in-tree with nvcc for sm_100a (B200), and the small CPython helper _pyingest.  No torch
involved: the library only needs the CUDA runtime (linked statically) and exposes the C-ABI of include/distgcn_b200.h.
Translation units are compiled in parallel into distgcn_b200/build/*.o and re-used while their sources are unchanged."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB_NAME = "libdistgcn_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["dg_api.cu", "dg_gcn.cu", "dg_lgs.cu", "dg_fused.cu", "dg_tc.cu", "dg_ingest.cu", "dg_stream.cu", "dg_wireless.cu"]
HEADERS = ["dg_common.cuh", os.path.join("..", "..", "include", "distgcn_b200.h")]
PYINGEST_SRC = os.path.join(CSRC, "pyingest.c")
PYINGEST_PATH = os.path.join(HERE, "_pyingest" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]
LINK_FLAGS = ["--shared", "-cudart", "static"]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def _newer(target: str, deps) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    deps = [os.path.join(CSRC, s) for s in _sources() + HEADERS] + [os.path.abspath(__file__)]
    return _newer(LIB_PATH, deps) or _newer(PYINGEST_PATH, [PYINGEST_SRC])


def build_pyingest(verbose: bool = False) -> str:
    """The CPython helper that walks a list of scipy matrices in C (csrc/pyingest.c): plain gcc against Python.h."""
    if not _newer(PYINGEST_PATH, [PYINGEST_SRC]):
        return PYINGEST_PATH
    cc = os.environ.get("CC") or shutil.which("gcc") or shutil.which("cc")
    if not cc:
        raise RuntimeError("no C compiler found for _pyingest")
    cmd = [cc, "-O2", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"], PYINGEST_SRC, "-o", PYINGEST_PATH]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return PYINGEST_PATH


def build_library(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    if not force and not needs_build():
        return LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    objs = []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or extra_flags or _newer(obj, [src] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    run([nvcc] + NVCC_FLAGS + LINK_FLAGS + ["-o", LIB_PATH] + objs)
    if os.path.isfile(PYINGEST_SRC):
        build_pyingest(verbose)
    return LIB_PATH


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True, extra_flags=[a for a in sys.argv[1:] if a != "--force"])
    print(LIB_PATH)
