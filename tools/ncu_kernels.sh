#!/bin/bash
# ncu --set full of selected kernels of tools/one_forward.py:  ncu_kernels.sh <name> <regex> <skip> <count>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 \
    -o gpurun_out/$1 -f python tools/one_forward.py S 32 > gpurun_out/$1.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
ls -la gpurun_out/ | head; du -sh gpurun_out
