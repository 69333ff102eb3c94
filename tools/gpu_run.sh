mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>gpurun_out/ts.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['roofline']['kernel'][:12])"
tail -3 gpurun_out/ts.err
DG_FUSED_TIMING=1 DG_TC_TILE_DUMP=gpurun_out/tiles_ts.txt timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/timing_ts.err >/dev/null
grep "tc t" gpurun_out/timing_ts.err | tail -13
