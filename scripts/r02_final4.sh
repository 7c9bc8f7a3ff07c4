#!/bin/bash
# final check of the tree after the STFT experiment was backed out: GPU tests + smoke
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 ) > gpurun_out/r02_pytest_final4.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/r02_smoke4.log 2>&1
( timeout 120 python scripts/r02_stft_time.py 2>&1 | tail -1 ) > gpurun_out/r02_stft_final4.log
grep -E "passed|failed" gpurun_out/r02_pytest_final4.log | tail -1; grep -E "^E |FAILED" gpurun_out/r02_pytest_final4.log | head -5 | cut -c1-250
head -1 gpurun_out/r02_smoke4.log; cat gpurun_out/r02_stft_final4.log
