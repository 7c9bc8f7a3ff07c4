#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_pytest_final.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/r02_smoke.log 2>&1
( time timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/r02_bench_full_n1.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -3 ) > gpurun_out/r02_bench_ref.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_final.log | tail -1; cat gpurun_out/r02_smoke.log | head -3; grep "^{" gpurun_out/r02_bench_full_n1.log | cut -c1-200; grep real gpurun_out/r02_bench_full_n1.log; grep "^{" gpurun_out/r02_bench_ref.log | cut -c1-200
