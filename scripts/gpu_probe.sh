#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu4.log 2>&1
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r01_tc2.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch3.log 2>&1
tail -8 gpurun_out/pytest_gpu4.log; tail -3 gpurun_out/bench4.log | cut -c1-400
