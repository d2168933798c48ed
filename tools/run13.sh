#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_hie_modules.py tests/test_gpu_parity_full_dims.py tests/test_gpu_solver_loop.py -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -40 > gpurun_out/r2_tests13.log
tail -8 gpurun_out/r2_tests13.log
timeout 400 python bench.py --config c3 --steps 20 --warmup 5 > gpurun_out/r2j_bench_c3.json 2> gpurun_out/r2j_bench_c3.err; echo "c3 rc=$?"
