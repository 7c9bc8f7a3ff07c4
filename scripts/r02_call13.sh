#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gan_gpu.py tests/test_librivox_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest13.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench13.log 2>&1
( VIAI_BATCHED_PACK=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench13_nobatch.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest13.log | tail -2; grep -E "^E |FAILED" gpurun_out/r02_pytest13.log | head -5 | cut -c1-250
for f in gpurun_out/r02_bench13.log gpurun_out/r02_bench13_nobatch.log; do grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"launches_per_step": [0-9]*' $f | head -1; done
