#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -30 > gpurun_out/r2_tests25.log
tail -5 gpurun_out/r2_tests25.log
for c in c1 c4; do
timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2q_bench_$c.json 2> gpurun_out/r2q_bench_$c.err; echo "bench $c rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2q_bench_$c.json').read());print('value %.0f ms %.3f e2e %.0f launches/step %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']/d['steps']))"
done
