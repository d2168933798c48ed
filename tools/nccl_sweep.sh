#!/bin/bash
# c2 at N GPUs under different NCCL CTA caps: value, ms/step and the exposed time of the exchange (bench.py's dry-run leg)
n=${1:-2}; shift
mkdir -p gpurun_out
for cap in ${@:-default 4 8 16}; do
  unset NCCL_MAX_CTAS NCCL_MIN_CTAS
  case "$cap" in
    default) ;;
    min*) export NCCL_MIN_CTAS=${cap#min} ;;      # "min32": at least 32 CTAs per collective
    *) export NCCL_MAX_CTAS=$cap ;;
  esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
      bench.py --config c2 --gpus $n --steps 20 --warmup 5 > gpurun_out/sweep_$cap.json 2> gpurun_out/sweep_$cap.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/sweep_$cap.json').read().strip().splitlines()[-1])
    c = d.get('gradient_exchange_exposed') or {}
    print('NCCL CTAs: $cap  value %.0f  %.3f ms  dry %.3f  exposed %.3f  e2e %.0f' % (d['value'], d['ms_per_step'],
          c.get('ms_per_step_without_collectives', 0), c.get('exposed_comm_ms', 0), d['e2e']['value']))
except Exception as e:
    print('$cap failed', e); print(open('gpurun_out/sweep_$cap.err').read()[-800:])
PY
done
