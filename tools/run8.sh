#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 bench.py --config c2 --gpus 2 --steps 20 --warmup 5 --shard 1 > gpurun_out/r2_n2_shard1.json 2> gpurun_out/r2_n2_shard1.err
echo "n2 shard=1 rc=$?"; grep -v "Warning\|kl_div" gpurun_out/r2_n2_shard1.err | tail -6
