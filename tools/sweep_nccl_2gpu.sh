for cfg in "0 1" "16 1" "8 1" "0 0" "16 0"; do
  set -- $cfg
  if [ "$1" != "0" ]; then export NCCL_MAX_CTAS=$1; else unset NCCL_MAX_CTAS; fi
  export VQA_B200_DDP_DEFER=$2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/sweep2_$1_$2.json 2> gpurun_out/sweep2_$1_$2.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/sweep2_$1_$2.json'))
print('max_ctas=$1 defer=$2', round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
"
done
