#!/bin/bash
# round 2, call 14: alternating sweep direction of the norm passes (VIAI_NORM_WALK) + statistics fused into the Cin=1 convolution
# (VIAI_CIN1_STATS): GPU tests, then A/B/C bench of the C2 step, then the per-operator table
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_pytest14.log 2>&1
B="python bench.py --steps 30 --no-cpu-baseline --no-wavenet --no-extra"
( timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench14_default.log 2>&1
( VIAI_NORM_WALK=0 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench14_nowalk.log 2>&1
( VIAI_CIN1_STATS=0 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench14_nocin1.log 2>&1
( VIAI_NORM_WALK=0 VIAI_CIN1_STATS=0 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench14_old.log 2>&1
( timeout 200 python scripts/r02_op_table.py 2>&1 | tail -130 ) > gpurun_out/r02_op_table14.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest14.log | tail -2; grep -E "^E |FAILED" gpurun_out/r02_pytest14.log | head -8 | cut -c1-250
for f in default nowalk nocin1 old; do echo -n "$f: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench14_$f.log | head -1; grep -o '"launches_per_step": [0-9]*' gpurun_out/r02_bench14_$f.log | head -1; done
head -3 gpurun_out/r02_op_table14.log; tail -1 gpurun_out/r02_op_table14.log
