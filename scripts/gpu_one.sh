#!/bin/bash
# one ncu --set full capture of the SKIP-th conv_tc launch of an eager step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k "regex:${KERNEL:-conv_tc_kernel}" --launch-skip ${SKIP:-19} --launch-count 1 -o gpurun_out/prof_one -f \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log | cut -c1-200
