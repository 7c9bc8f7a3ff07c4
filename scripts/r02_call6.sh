#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest6.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench6.log 2>&1
for L in block5 d_conv2_1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kernel" -s 3 -c 3 -o gpurun_out/r02_prof_$L -f python scripts/r02_layer_probe.py $L 2 > gpurun_out/r02_prof_$L.log 2>&1
done
tail -3 gpurun_out/r02_pytest6.log; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench6.log | head -1; ls -la gpurun_out/*.ncu-rep
