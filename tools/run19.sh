#!/bin/bash
mkdir -p gpurun_out
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=29700
for c in "$@"; do
p=$((p+1))
timeout 300 $TR --master-port $p bench.py --config $c --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_${c}_n${N}.json 2> gpurun_out/r2_${c}_n${N}.err; echo "$c n$N rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/r2_${c}_n${N}.json').read());print('$c n$N value %.0f ms %.3f e2e %.0f'%(d['value'],d['ms_per_step'],d['e2e']['value']))"
done
