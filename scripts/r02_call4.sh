#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_table.csv gpurun_out/parity_table.csv.flips
export VIAI_PARITY_TABLE=$PWD/gpurun_out/parity_table.csv
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -60 ) > gpurun_out/r02_pytest4.log 2>&1
unset VIAI_PARITY_TABLE
( time timeout 900 python bench.py 2>&1 | tail -5 ) > gpurun_out/r02_bench_full_n1.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 ) > gpurun_out/r02_bench_ref.log 2>&1
tail -3 gpurun_out/r02_pytest4.log; tail -4 gpurun_out/r02_bench_full_n1.log | cut -c1-1500; tail -4 gpurun_out/r02_bench_ref.log | cut -c1-600
