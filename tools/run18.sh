#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
VQA_B200_BENCH_FORCE_CAPTURE_FAILURE=1 timeout 200 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_fallback_n1.json 2> gpurun_out/r2_fallback_n1.err; echo "fallback n1 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2_fallback_n1.json').read());print('n1 value %.0f graph: %s'%(d['value'],d['config']['cuda_graph']))"
VQA_B200_BENCH_FORCE_CAPTURE_FAILURE=1 timeout 200 $TR --master-port 29601 bench.py --config c2 --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_fallback_n2.json 2> gpurun_out/r2_fallback_n2.err; echo "fallback n2 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2_fallback_n2.json').read());print('n2 value %.0f graph: %s'%(d['value'],d['config']['cuda_graph']))"
grep -v "Warning\|kl_div" gpurun_out/r2_fallback_n2.err | tail -5
for c in c1 c4; do
timeout 300 $TR --master-port 2961${c:1} bench.py --config $c --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_${c}_n2.json 2> gpurun_out/r2_${c}_n2.err; echo "$c n2 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2_${c}_n2.json').read());print('$c n2 value %.0f ms %.3f e2e %.0f | %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['config']['gradient_exchange'][:90]))"
done
