#!/bin/bash
mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29561 bench.py --config c2 --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_n${N}_shard1.json 2> gpurun_out/r2_n${N}_shard1.err
echo "n$N rc=$?"; grep -v "Warning\|kl_div" gpurun_out/r2_n${N}_shard1.err | tail -6
