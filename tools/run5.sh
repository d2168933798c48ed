#!/bin/bash
mkdir -p gpurun_out
for g in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --config c2 --gpus 2 --steps 20 --warmup 5 --graph $g > gpurun_out/r2_n2_graph$g.json 2> gpurun_out/r2_n2_graph$g.err
echo "n2 graph=$g rc=$?"; tail -4 gpurun_out/r2_n2_graph$g.err
done
