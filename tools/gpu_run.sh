mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 20 2>gpurun_out/two.err > gpurun_out/s5e_bench_ba500.json; python -c "
import json; d=json.load(open('gpurun_out/s5e_bench_ba500.json')); print('bench', d['value'], d['ms_per_step'], d['roofline']['avg_launch_us'], d['roofline']['share_of_step'], d['roofline']['frac'], d['roofline']['measured_on'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['gpu_launches'], d['cpu_baseline']['value'])"
tail -3 gpurun_out/two.err
timeout 300 python bench.py --workload er500 --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('er500', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['roofline']['kernel'][:20], d['roofline']['avg_launch_us'])"
