#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 tools/ddp_parity_n2.py > gpurun_out/r2_ddp_parity_n2.json 2> gpurun_out/r2_ddp_parity_n2.err; echo "parity rc=$?"; cat gpurun_out/r2_ddp_parity_n2.json
PROBE_LIMIT=50 timeout 80 $TR --master-port 29522 tools/ddp_graph_probe.py > gpurun_out/r2_probe_a.log 2>&1; echo "probe small rc=$?"; grep -v Warning gpurun_out/r2_probe_a.log | tail -30
