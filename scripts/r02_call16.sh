#!/bin/bash
# round 2, call 16: finer threshold scan of VIAI_NORM_WALK_MB (tensor sizes in the C2 step are 256 / 128 / 64 / 32 / 16 / ... MiB),
# wide pixels in flight in the thin weight-gradient kernel (VIAI_WGT_UN)
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_layers_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 ) > gpurun_out/r02_pytest16_un2.log 2>&1
( VIAI_WGT_UN=4 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_layers_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 ) > gpurun_out/r02_pytest16_un4.log 2>&1
tail -1 gpurun_out/r02_pytest16_un2.log; tail -1 gpurun_out/r02_pytest16_un4.log
B="python bench.py --steps 60 --no-cpu-baseline --no-wavenet --no-extra"
for mb in 48 30 14 6; do
  ( VIAI_NORM_WALK_MB=$mb timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench16_mb$mb.log 2>&1
  echo -n "mb$mb: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench16_mb$mb.log | head -1
done
( VIAI_WGT_UN=4 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench16_un4.log 2>&1
echo -n "un4 (mb48): "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench16_un4.log | head -1
( timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench16_mb48b.log 2>&1
echo -n "mb48 again: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench16_mb48b.log | head -1
