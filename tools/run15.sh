#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --config c2 --precision fp32 --steps 20 --warmup 5 > gpurun_out/r2l_bench_c2_fp32.json 2> gpurun_out/r2l_bench_c2_fp32.err; echo "c2 fp32 rc=$?"; tail -2 gpurun_out/r2l_bench_c2_fp32.err
for c in c1 c4 c5; do
timeout 400 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r2l_bench_$c.json 2> gpurun_out/r2l_bench_$c.err; echo "$c rc=$?"
done
timeout 300 python bench.py --config c1 --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_c1_ref.json 2>/dev/null; echo "c1 ref rc=$?"
for f in c2_fp32 c1 c4 c5 c1_ref; do python -c "
import json;d=json.loads(open('gpurun_out/r2l_bench_$f.json').read());print('$f value %.0f ms %.3f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"; done
