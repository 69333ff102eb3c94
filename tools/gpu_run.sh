python tools/host_cost.py 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or full_config or host or heuristics or dit" 2>&1 | tail -2
timeout 300 python bench.py --workload er500 --steps 100 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('er500', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['paths_agree'])"
