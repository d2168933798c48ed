#!/bin/bash
# round-2 GPU call 2: captured train iteration + feed
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_graph.py tests/test_feed.py tests/test_gpu_lstm.py tests/test_gpu_optim.py -q -m gpu -x 2>&1 | tail -40 > gpurun_out/r2_tests2a.log
tail -5 gpurun_out/r2_tests2a.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_tests2.log
tail -5 gpurun_out/r2_tests2.log
for c in c2 c1 c4 c3; do
  timeout 500 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2g_bench_${c}.json 2> gpurun_out/r2g_bench_${c}.err
  echo "bench $c rc=$?"; tail -3 gpurun_out/r2g_bench_${c}.err
done
