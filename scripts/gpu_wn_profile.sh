#!/bin/bash
# ncu launch list of eager WaveNet training steps (B=4, T=8000)
mkdir -p gpurun_out
cat > /tmp/wnp.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
print(bench.time_wavenet_train(torch, B=4, T=8000, steps=1, warmup=1, graph=False))
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_wn_train.csv python /tmp/wnp.py > gpurun_out/wn_profile.log 2>&1
tail -3 gpurun_out/wn_profile.log
wc -l gpurun_out/launches_wn_train.csv
