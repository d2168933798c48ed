#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
PARITY_SHARD=1 timeout 200 $TR --master-port 29531 tools/ddp_parity_n2.py > gpurun_out/r2_ddp_parity_n2_sharded.json 2> gpurun_out/r2_ddp_parity_n2_sharded.err; echo "parity sharded rc=$?"; cat gpurun_out/r2_ddp_parity_n2_sharded.json; grep -v Warning gpurun_out/r2_ddp_parity_n2_sharded.err | grep -i "error\|Traceback" -A8 | head -40
PARITY_SHARD=0 timeout 200 $TR --master-port 29532 tools/ddp_parity_n2.py > gpurun_out/r2_ddp_parity_n2.json 2> gpurun_out/r2_ddp_parity_n2.err; echo "parity rc=$?"; cat gpurun_out/r2_ddp_parity_n2.json
PROBE_SHARD=1 PROBE_LIMIT=50 timeout 80 $TR --master-port 29533 tools/ddp_graph_probe.py > gpurun_out/r2_probe_b.log 2>&1; echo "probe sharded+graph rc=$?"; grep -v Warning gpurun_out/r2_probe_b.log | tail -12
for sh in 1 0; do
timeout 300 $TR --master-port 2954$sh bench.py --config c2 --gpus 2 --steps 20 --warmup 5 --shard $sh > gpurun_out/r2_n2_shard$sh.json 2> gpurun_out/r2_n2_shard$sh.err
echo "n2 shard=$sh rc=$?"; grep -v "Warning\|kl_div" gpurun_out/r2_n2_shard$sh.err | tail -4
done
