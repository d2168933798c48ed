#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_solver_loop.py tests/test_gpu_train_graph.py -q -m gpu --tb=short 2>&1 | grep -v Warning > gpurun_out/r2_tests4.log
tail -5 gpurun_out/r2_tests4.log
timeout 500 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
echo "bench c2 rc=$?"; tail -3 gpurun_out/r2f_bench_c2.err
