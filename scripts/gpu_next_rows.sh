#!/bin/bash
# new-row tests (all, no -x) + WaveNet training probe
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests/test_next_rows_gpu.py -q -p no:cacheprovider 2>&1 | tail -150 ) > gpurun_out/next_rows.log 2>&1
( time timeout 200 python scripts/wn_train_probe.py ) > gpurun_out/wn_train.log 2>&1
tail -5 gpurun_out/next_rows.log; tail -4 gpurun_out/wn_train.log
