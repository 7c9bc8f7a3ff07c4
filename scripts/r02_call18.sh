#!/bin/bash
# round 2, call 18: __launch_bounds__(256, 3) on the fused-statistics 1x4 Cin=1 kernel (80 registers, 3 CTAs per SM) vs (256, 1)
mkdir -p gpurun_out
B="python bench.py --steps 60 --no-cpu-baseline --no-wavenet --no-extra"
( timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -p no:cacheprovider -k "cin1 or conv2d" 2>&1 | tail -2 ) > gpurun_out/r02_pytest18.log 2>&1
tail -1 gpurun_out/r02_pytest18.log
( timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench18_lb3.log 2>&1
echo -n "bounds (256,3): "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench18_lb3.log | head -1
sed -i 's/__launch_bounds__(256, (STATS \&\& R_ \* S_ <= 4) ? 3 : 1) conv_cin1_kernel/__launch_bounds__(256) conv_cin1_kernel/' vision-infused-audio-inpainter-viai_b200/csrc/conv_thin.cu
grep -c "__launch_bounds__(256) conv_cin1_kernel" vision-infused-audio-inpainter-viai_b200/csrc/conv_thin.cu
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
( timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench18_lb1.log 2>&1
echo -n "bounds (256,1): "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench18_lb1.log | head -1
