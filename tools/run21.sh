#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_graph.py tests/test_gpu_parity_full_dims.py tests/test_gpu_optim.py -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -30 > gpurun_out/r2_tests21.log
tail -6 gpurun_out/r2_tests21.log
