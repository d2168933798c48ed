#!/bin/bash
# round-2 GPU call 3: fused block (batched projections), solver-loop / DataParallel / feed tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -60 > gpurun_out/r2_tests3.log
tail -12 gpurun_out/r2_tests3.log
timeout 500 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_c2.json 2> gpurun_out/r2f_bench_c2.err
echo "bench c2 rc=$?"; tail -3 gpurun_out/r2f_bench_c2.err
