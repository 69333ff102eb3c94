mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['roofline']['kernel'][:12])"
DG_FUSED_TIMING=1 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/timing_v4.err >/dev/null
grep "tc t" gpurun_out/timing_v4.err | tail -12
