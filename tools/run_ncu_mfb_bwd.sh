#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/mfb_bwd_probe.py single thread && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:mfb_bwd -s 1 -c 1 -o gpurun_out/r02_mfb_bwd_thread python tools/mfb_bwd_probe.py single thread > gpurun_out/ncu_mb.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_mb.log; ls -la gpurun_out/r02_mfb_bwd_thread*
