#!/bin/bash
# final validation of round 2 (after the norm-sweep / fused Cin=1 statistics work): GPU tests, smoke, bench N=1, ncu launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_pytest_final3.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/r02_smoke3.log 2>&1
( time timeout 900 python bench.py 2>&1 | tail -3 ) > gpurun_out/r02_bench_full3_n1.log 2>&1
( timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1000 --csv --log-file gpurun_out/r02_launches_final3.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet --no-extra > gpurun_out/r02_ncu_bench_final3.log 2>&1 )
python scripts/summarize_launches.py gpurun_out/r02_launches_final3.csv --last-step adam_kernel 2 > gpurun_out/r02_final3_launch_shares.csv 2> gpurun_out/r02_final3_launch_shares.err
grep -E "passed|failed" gpurun_out/r02_pytest_final3.log | tail -1; grep -E "^E |FAILED" gpurun_out/r02_pytest_final3.log | head -5 | cut -c1-250
head -3 gpurun_out/r02_smoke3.log; grep "^{" gpurun_out/r02_bench_full3_n1.log | cut -c1-300; grep real gpurun_out/r02_bench_full3_n1.log
head -8 gpurun_out/r02_final3_launch_shares.csv; tail -1 gpurun_out/r02_final3_launch_shares.csv; cat gpurun_out/r02_final3_launch_shares.err | tail -2
