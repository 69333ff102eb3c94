timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core or full_config or zero_weight or solve_membership" 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['roofline']['avg_launch_us'], d['e2e']['ms_per_step'], d['config']['paths_agree'])"
done
DG_FUSED_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 >/dev/null | grep "tc timing\] greedy\|tc timing\] total" | tail -2
