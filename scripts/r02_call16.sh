#!/bin/bash
# round 2, call 16: finer threshold scan of VIAI_NORM_WALK_MB (tensor sizes in the C2 step are 256 / 128 / 64 / 32 / 16 / ... MiB)
mkdir -p gpurun_out
B="python bench.py --steps 60 --no-cpu-baseline --no-wavenet --no-extra"
for mb in 60 30 14 6 60; do
  ( VIAI_NORM_WALK_MB=$mb timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench16_mb$mb.log 2>&1
  echo -n "mb$mb: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench16_mb$mb.log | head -1
done
