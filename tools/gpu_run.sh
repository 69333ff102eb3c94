mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['roofline']['avg_launch_us'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['one_call_at_a_time']['ms_per_step'], d['config']['paths_agree'], d['gpu_launches'], d['e2e']['gpu_launches'])"
