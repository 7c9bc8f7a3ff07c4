#!/bin/bash
# ncu --set full capture of selected kernels (regex in $1, count in $2) from one eager GAN step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k "regex:$1" -c ${2:-20} -o gpurun_out/prof_sel -f \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet > gpurun_out/ncu_sel.log 2>&1
tail -3 gpurun_out/ncu_sel.log | cut -c1-300
