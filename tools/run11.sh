#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -90 > gpurun_out/r2_tests11.log
tail -14 gpurun_out/r2_tests11.log
for c in c2 c3 c5; do
timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2h_bench_$c.json 2> gpurun_out/r2h_bench_$c.err; echo "$c rc=$?"
done
