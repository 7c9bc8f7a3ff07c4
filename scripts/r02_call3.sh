#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.csv gpurun_out/parity_table.csv.flips
export VIAI_PARITY_TABLE=$PWD/gpurun_out/parity_table.csv
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 ) > gpurun_out/r02_pytest3_all.log 2>&1
unset VIAI_PARITY_TABLE
for cfg in "fp16x3 x3 tf32r" "fp16x3 tf32x3 tf32r" "fp16x3 x3 fp32" "fp16x3 tf32x3 fp32" "bf16x3 x3 tf32r" "tf32x3 x3 tf32r"; do
  set -- $cfg
  ( VIAI_PRECISION=$1 VIAI_DGRAD=$2 VIAI_WGRAD=$3 timeout 300 python scripts/r02_parity_components.py 2>&1 | grep "^precision" ) >> gpurun_out/r02_parity_components.log 2>&1
done
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet 2>&1 | tail -3 ) > gpurun_out/r02_bench_c3.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest3_all.log | tail -3; cat gpurun_out/r02_parity_components.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_c3.log | head -1
