#!/bin/bash
# final validation of round 2 (after the WaveNet kernel work): GPU tests, smoke, bench N=1
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_pytest_final2.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/r02_smoke2.log 2>&1
( time timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/r02_bench_full2_n1.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_final2.log | tail -1; head -3 gpurun_out/r02_smoke2.log; grep "^{" gpurun_out/r02_bench_full2_n1.log | cut -c1-300; grep real gpurun_out/r02_bench_full2_n1.log
