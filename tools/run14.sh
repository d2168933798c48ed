#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -60 > gpurun_out/r2_tests14.log
tail -8 gpurun_out/r2_tests14.log
timeout 400 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "c2 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2k_bench_c2.json').read());print('c2 value %.0f ms %.3f e2e %.0f host_enq %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_enqueue_ms_per_step']))"
timeout 400 python bench.py --config c2 --steps 20 --warmup 5 --graph 0 --no-cpu-baseline > gpurun_out/r2k_bench_c2_eager.json 2> gpurun_out/r2k_bench_c2_eager.err; echo "c2 eager rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2k_bench_c2_eager.json').read());print('c2 eager value %.0f ms %.3f host_enq %.3f'%(d['value'],d['ms_per_step'],d['host_enqueue_ms_per_step']))"
