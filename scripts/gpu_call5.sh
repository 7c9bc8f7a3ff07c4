#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench1.log 2>&1
( VIAI_TC_SA=3 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_sa3.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet > gpurun_out/ncu_launch.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench1.log
