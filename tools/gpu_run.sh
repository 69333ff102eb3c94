mkdir -p gpurun_out
timeout 300 python bench.py --workload synth-er-16384 --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/synth.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('synth16k', d['ms_per_step'], d['wall_ms_per_step'], d['roofline']['avg_launch_us'], d['roofline']['kernel'][:20], d['gpu_launches'], d['e2e']['ms_per_step'], d['config']['paths_agree'])"
tail -3 gpurun_out/synth.err
DG_FUSED_TIMING=1 timeout 300 python bench.py --workload synth-er-16384 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "tc t" | head -14
