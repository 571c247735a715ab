#!/bin/bash
# ncu --set full of selected kernels of one training step: ncu_train.sh <name> <kernel regex> <skip> <count> [model] [batch]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 \
    -o gpurun_out/$1 -f python tools/train_step_bench.py ${5:-S} ${6:-16} auto 1 > gpurun_out/$1.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/$1.raw.csv
