#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest9.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench9.log 2>&1
( timeout 300 python -m pytest tests/test_stft_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3 ) > gpurun_out/r02_stft_test9.log 2>&1
( timeout 300 python -c "
import sys, json; sys.path.insert(0, \".\")
import torch, bench
print(json.dumps(bench.run_stft(bench.Ctx(torch, None, 0, 1), bench.peaks())))" 2>&1 | tail -1 ) > gpurun_out/r02_stft_bench9.log 2>&1
( timeout 600 python scripts/c3_probe.py 2>&1 | tail -2 ) > gpurun_out/r02_c3_probe.log 2>&1
( timeout 600 python scripts/c3_kernels.py 2>&1 | tail -60 ) > gpurun_out/r02_c3_kernels.csv 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest9.log | tail -2; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench9.log | head -1; cat gpurun_out/r02_c3_probe.log | cut -c1-400; head -14 gpurun_out/r02_c3_kernels.csv
cat gpurun_out/r02_stft_test9.log; cut -c1-300 gpurun_out/r02_stft_bench9.log
