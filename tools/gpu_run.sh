for i in 1 2 3; do
timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['roofline']['avg_launch_us'], d['e2e']['ms_per_step'], d['config']['paths_agree'])"
done
