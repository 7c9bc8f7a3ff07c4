#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest9.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench9.log 2>&1
( timeout 600 python scripts/c3_probe.py 2>&1 | tail -2 ) > gpurun_out/r02_c3_probe.log 2>&1
( timeout 600 python scripts/c3_kernels.py 2>&1 | tail -60 ) > gpurun_out/r02_c3_kernels.csv 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest9.log | tail -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench9.log | head -1; cat gpurun_out/r02_c3_probe.log | cut -c1-400; head -14 gpurun_out/r02_c3_kernels.csv
