#!/bin/bash
# records of the final library: per-operator table of one eager C2 step, C3 probe at the per-GPU share of the 8-GPU run
mkdir -p gpurun_out
( timeout 200 python scripts/r02_op_table.py 2>&1 | tail -130 ) > gpurun_out/r02_op_table_final.log 2>&1
( timeout 300 python scripts/c3_probe.py 2>&1 | tail -2 ) > gpurun_out/r02_c3_probe_final.log 2>&1
head -2 gpurun_out/r02_op_table_final.log | cut -c1-200; tail -1 gpurun_out/r02_op_table_final.log; cat gpurun_out/r02_c3_probe_final.log | cut -c1-400
