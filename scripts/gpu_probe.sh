#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python scripts/tc_probe.py 4 2>&1 | tail -16 ) > gpurun_out/probe5.log 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu3.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r01_tc.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
cat gpurun_out/probe5.log; tail -15 gpurun_out/pytest_gpu3.log; tail -3 gpurun_out/bench3.log | cut -c1-900
