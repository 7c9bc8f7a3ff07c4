#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -s 0 -c 7 -o gpurun_out/r02_prof_d_conv3 -f python scripts/r02_layer_probe.py d_conv3 1 > gpurun_out/r02_prof_d_conv3.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1530 -c 520 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet --no-extra > gpurun_out/r02_ncu_bench_final.log 2>&1 )
ls -la gpurun_out/r02_prof_d_conv3.ncu-rep; wc -l gpurun_out/r02_launches_final.csv
