#!/bin/bash
# round-2 ncu evidence: launch list of two eager train steps + full captures of the GEMM / pooling kernels
mkdir -p gpurun_out
export VQA_B200_LSTM_COOP=0
CMD="python bench.py --config c2 --steps 2 --warmup 3 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 330 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel|softmax_pool" -s 60 -c 34 -o gpurun_out/r02_prof $CMD > gpurun_out/ncu_f.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/r02_prof* gpurun_out/r02_launches.csv
