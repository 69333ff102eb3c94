for u in 2 8; do
DISTGCN_B200_LIB=$PWD/distgcn_b200/libdg_u$u.so timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unroll $u', d['ms_per_step'], d['roofline']['avg_launch_us'], d['config']['paths_agree'])"
done
