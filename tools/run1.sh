#!/bin/bash
# round-2 GPU call 1: full test suite (new full-dimension parity tests), then every bench configuration once
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rA 2>&1 | grep -v "^PASSED" | tail -60 > gpurun_out/r2_tests1.log
for c in c2 c1 c3 c4 c5; do
  timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2_bench_${c}.json 2> gpurun_out/r2_bench_${c}.err
  echo "bench $c rc=$?"
done
timeout 300 python bench.py --config c2 --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c2_ref.json 2> gpurun_out/r2_bench_c2_ref.err
tail -5 gpurun_out/r2_tests1.log
