#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_stft_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_stft_test.log 2>&1
( timeout 300 python scripts/r02_op_table.py 2>&1 | tail -90 ) > gpurun_out/r02_op_table.log 2>&1
( timeout 300 python - <<'PY' 2>&1 | tail -5
import sys, json
sys.path.insert(0, '.')
import torch, bench
class C: pass
ctx = bench.Ctx(torch, None, 0, 1)
print(json.dumps(bench.run_stft(ctx, bench.peaks())))
import os
os.environ["VIAI_STFT_FAST"] = "0"
PY
) > gpurun_out/r02_stft_bench.log 2>&1
( VIAI_STFT_FAST=0 timeout 300 python -c "
import sys, json
sys.path.insert(0, '.')
import torch, bench
ctx = bench.Ctx(torch, None, 0, 1)
print(json.dumps(bench.run_stft(ctx, bench.peaks())))" 2>&1 | tail -2 ) > gpurun_out/r02_stft_bench_old.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet --no-extra > gpurun_out/r02_ncu_bench.log 2>&1 )
tail -3 gpurun_out/r02_stft_test.log; cat gpurun_out/r02_stft_bench.log | cut -c1-600; cat gpurun_out/r02_stft_bench_old.log | cut -c1-300; head -5 gpurun_out/r02_op_table.log; wc -l gpurun_out/r02_launches.csv
