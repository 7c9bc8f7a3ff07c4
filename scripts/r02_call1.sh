#!/bin/bash
# round 2, GPU call 1: the whole GPU suite on the benched (default) precision + the per-tensor parity table; the opt-in stem
# kernels; cost of the 3-term data gradient and of the operand rounding in the weight gradient.
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.csv
export VIAI_PARITY_TABLE=$PWD/gpurun_out/parity_table.csv
( time timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -150 ) > gpurun_out/r02_pytest1.log 2>&1
( time timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -250 ) > gpurun_out/r02_pytest1_all.log 2>&1
unset VIAI_PARITY_TABLE
( VIAI_FAST_STEM=1 timeout 300 python -m pytest tests/test_fast_stem_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > gpurun_out/r02_fast_stem.log 2>&1
for cfg in "x3 1" "tf32 1" "x3 0" "tf32 0"; do
  set -- $cfg
  ( VIAI_DGRAD=$1 VIAI_WGRAD_ROUND=$2 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet 2>&1 | tail -3 ) > gpurun_out/r02_bench_dgrad_$1_round_$2.log 2>&1
done
tail -3 gpurun_out/r02_pytest1.log; tail -3 gpurun_out/r02_pytest1_all.log; tail -3 gpurun_out/r02_fast_stem.log
for f in gpurun_out/r02_bench_dgrad_*; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; done
