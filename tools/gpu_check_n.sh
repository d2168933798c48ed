#!/bin/bash
# N GPUs of one box: bench lines of the given configurations under torchrun (what the round-end scaling run does).
#   gpurun --gpus N --timeout 1800 -- 'bash tools/gpu_check_n.sh N [configs...]'
n=${1:-2}; shift
mkdir -p gpurun_out
for c in ${@:-c2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --config $c --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${c}_n$n.json 2> gpurun_out/bench_${c}_n$n.err
  echo "bench $c N=$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_${c}_n$n.json').read().strip().splitlines()[-1])
    print('$c N=$n value %.0f %s  %.3f ms  e2e %.0f  graph: %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'],
          str(d['config'].get('cuda_graph'))[:60]))
except Exception as e:
    print('no line:', e); print(open('gpurun_out/bench_${c}_n$n.err').read()[-1500:])
PY
done
