#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
PARITY_SHARD=1 timeout 200 $TR --master-port 29571 tools/ddp_parity_n2.py 2> gpurun_out/p1.err | cut -c1-420; echo "parity sharded rc=$?"
PARITY_SHARD=0 timeout 200 $TR --master-port 29572 tools/ddp_parity_n2.py 2> gpurun_out/p0.err | cut -c1-420; echo "parity rc=$?"
for e in 1 0; do
VQA_B200_DDP_EARLY=$e timeout 300 $TR --master-port 2958$e bench.py --config c2 --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_n2_early$e.json 2> gpurun_out/r2_n2_early$e.err
echo "n2 early=$e rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r2_n2_early$e.json').read());print('value %.0f ms %.3f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
done
