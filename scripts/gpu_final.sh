#!/bin/bash
# round-end style check: every GPU test, smoke(), per-kernel table of the WaveNet training step, the bench line
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -80 ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( time timeout 300 python scripts/wn_train_kernels.py ) > gpurun_out/wn_train_kernels.csv 2> gpurun_out/wn_train_kernels.err
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; head -12 gpurun_out/wn_train_kernels.csv; tail -2 gpurun_out/bench_n1.log | cut -c1-400
