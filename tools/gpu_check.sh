#!/bin/bash
# One B200: smoke, the -m gpu suite and a bench line per configuration (what the round-end driver runs, plus c1/c3/c4/c5).
#   gpurun --timeout 2400 -- 'bash tools/gpu_check.sh [configs...]'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "Warning\|warnings.warn\|kl_div" | tail -30 > gpurun_out/tests_gpu.log
tail -3 gpurun_out/tests_gpu.log
for c in ${@:-c2}; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$c.json').read())
print('$c value %.0f %s  %.3f ms  e2e %.0f  launches/step %d  roofline %.3f' % (
    d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'] / d['steps'], d['roofline']['frac']))
PY
done
