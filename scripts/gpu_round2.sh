#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 200 python scripts/wn_train_probe.py ) > gpurun_out/wn_train.log 2>&1
( time timeout 400 python bench.py ) > gpurun_out/bench_n1.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; grep "^{" gpurun_out/wn_train.log | cut -c1-330; tail -4 gpurun_out/bench_n1.log | cut -c1-600
