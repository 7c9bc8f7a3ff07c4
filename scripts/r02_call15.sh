#!/bin/bash
# round 2, call 15: size threshold of the alternating norm sweeps (VIAI_NORM_WALK_MB), C2 step, CUDA-graph replay
mkdir -p gpurun_out
B="python bench.py --steps 60 --no-cpu-baseline --no-wavenet --no-extra"
for mb in 100 0 200 60; do
  ( VIAI_NORM_WALK_MB=$mb timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench15_mb$mb.log 2>&1
done
( VIAI_NORM_WALK=0 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench15_off.log 2>&1
( VIAI_NORM_WALK_MB=100 timeout 200 $B 2>&1 | tail -1 ) > gpurun_out/r02_bench15_mb100b.log 2>&1
( timeout 200 python scripts/r02_op_table.py 2>&1 | tail -130 ) > gpurun_out/r02_op_table15.log 2>&1
for f in mb100 mb0 mb200 mb60 off mb100b; do echo -n "$f: "; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench15_$f.log | head -1; done
tail -1 gpurun_out/r02_op_table15.log
