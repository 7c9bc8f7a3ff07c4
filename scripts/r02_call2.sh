#!/bin/bash
# round 2, GPU call 2: fp16x3 forward + decision-pattern gradient gate; full suite; step time of the new default
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.csv gpurun_out/parity_table.csv.flips
export VIAI_PARITY_TABLE=$PWD/gpurun_out/parity_table.csv
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -300 ) > gpurun_out/r02_pytest2_all.log 2>&1
unset VIAI_PARITY_TABLE
for prec in fp16x3 bf16x3; do
  ( VIAI_PRECISION=$prec timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet 2>&1 | tail -3 ) > gpurun_out/r02_bench_$prec.log 2>&1
done
grep -E "passed|failed" gpurun_out/r02_pytest2_all.log | tail -3
for f in gpurun_out/r02_bench_fp16x3.log gpurun_out/r02_bench_bf16x3.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; done
