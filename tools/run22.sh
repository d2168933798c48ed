#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_feed.py tests/test_gpu_train_graph.py -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -30 > gpurun_out/r2_tests22.log
tail -6 gpurun_out/r2_tests22.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2o_bench_c2.json').read());print('value %.0f ms %.3f e2e %.0f block %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['hot_path_block']))"
