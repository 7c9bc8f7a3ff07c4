#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest8.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench8.log 2>&1
( VIAI_TC_XF2=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench8_noxf2.log 2>&1
( VIAI_FUSE_BWD_REDUCE=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench8_nofuse.log 2>&1
( timeout 300 python scripts/r02_op_table.py 2>&1 | tail -95 ) > gpurun_out/r02_op_table8.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest8.log | tail -2; for f in gpurun_out/r02_bench8.log gpurun_out/r02_bench8_noxf2.log gpurun_out/r02_bench8_nofuse.log; do grep -o '"ms_per_step": [0-9.]*' $f | head -1; done; grep -E "convT 32->32" gpurun_out/r02_op_table8.log
