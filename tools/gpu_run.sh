mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -5
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/s5f_bench_ba500.json 2>gpurun_out/s5f.err; python -c "
import json; d=json.load(open('gpurun_out/s5f_bench_ba500.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['cpu_baseline']['value'], d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_tc3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python profiles/micro/ncu_summary.py gpurun_out/r01_tc3_launches.csv
ncu --set full --clock-control none --import-source on -k regex:tc_solve -s 2 -c 1 -f -o gpurun_out/r01_tc3_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
ls -la gpurun_out/r01_tc3_full.ncu-rep
