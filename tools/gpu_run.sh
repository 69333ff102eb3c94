mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dist_greedy or heuristics_entry" 2>&1 | tail -15
