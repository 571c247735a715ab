#!/bin/bash
# quick GPU visit: selected tests (-k "$1"), bench line, launch list of one bench step
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cut -c1-330 gpurun_out/bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
