#!/bin/bash
# launch list (gpu__time_duration per kernel) of eager train steps of one bench config:  run_ncu_launches.sh c4 [max launches]
cfg=${1:-c2}; cnt=${2:-1500}
mkdir -p gpurun_out
export VQA_B200_LSTM_COOP=0
CMD="python bench.py --config $cfg --steps 2 --warmup 3 --graph 0 --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain_$cfg.json 2> gpurun_out/ncu_plain_$cfg.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c $cnt --csv --log-file gpurun_out/launches_$cfg.csv $CMD > gpurun_out/ncu_l_$cfg.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_$cfg.csv
