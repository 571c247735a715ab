#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
if [ "$1" != "quick" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/launches.csv
fi
