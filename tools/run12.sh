#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_inference.py tests/test_gpu_train_graph.py tests/test_gpu_parity_hie_modules.py tests/test_gpu_parity.py -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -30 > gpurun_out/r2_tests12.log
tail -6 gpurun_out/r2_tests12.log
timeout 400 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo "c2 rc=$?"
