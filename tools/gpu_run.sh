mkdir -p gpurun_out
for k in 0 150 200 250; do
  export DG_TC_NOALIAS_MAX_NV=$k
  python bench.py --no-cpu-baseline --steps 100 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('noalias<=$k', d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['paths_agree'], d['roofline']['kernel'][:12])"
  DG_FUSED_TIMING=1 DG_TC_TILE_DUMP=gpurun_out/tiles_na$k.txt python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/timing_na$k.err >/dev/null
  grep "tc t" gpurun_out/timing_na$k.err | tail -3
done
DG_TC_NOALIAS_MAX_NV=250 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
