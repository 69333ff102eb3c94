mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) 2>&1 | tail -6
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/s5c_bench_ba500.json 2>gpurun_out/s5c.err; cut -c1-330 gpurun_out/s5c_bench_ba500.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s5c_bench_reference.json 2>>gpurun_out/s5c.err; cut -c1-200 gpurun_out/s5c_bench_reference.json
