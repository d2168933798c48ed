#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --config c5 --steps 20 --warmup 5 > gpurun_out/r2m_bench_c5.json 2> gpurun_out/r2m_bench_c5.err; echo "c5 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2m_bench_c5.json').read());print('c5 value %.0f ms %.3f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
timeout 200 python tools/overlap_probe.py 2> gpurun_out/overlap_probe.err | tee gpurun_out/overlap_probe.json
