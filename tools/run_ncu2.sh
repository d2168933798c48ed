#!/bin/bash
# DRAM traffic of the region-pooling kernel INSIDE the step (caches not flushed between kernels): how much of the DRAM
# time in its window is write-back of the dirty lines the kernels in front of it left in L2?
mkdir -p gpurun_out
export VQA_B200_LSTM_COOP=0
CMD="python bench.py --config c2 --steps 2 --warmup 3 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/ncu2_plain.json 2> gpurun_out/ncu2_plain.err &&
ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
    -k regex:"softmax_pool_fwd|attn_logits_fwd" -s 8 -c 12 --csv --log-file gpurun_out/r02_pool_in_step.csv $CMD > gpurun_out/ncu2.log 2>&1
echo "rc=$?"
