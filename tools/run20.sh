#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -40 > gpurun_out/r2_tests20.log
tail -6 gpurun_out/r2_tests20.log
timeout 400 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo "default bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2n_bench_default.json').read());print('default value %.0f ms %.3f e2e %.0f steps %d warmup %d launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['steps'],d['warmup'],d['gpu_launches']))"
