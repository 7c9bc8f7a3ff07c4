#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python scripts/bench_extra.py 16000 ) > gpurun_out/bench_extra.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/pytest_gpu5.log 2>&1
cat gpurun_out/bench_extra.log; tail -8 gpurun_out/pytest_gpu5.log
