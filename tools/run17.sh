#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -40 > gpurun_out/r2_tests17.log
tail -5 gpurun_out/r2_tests17.log
