#!/bin/bash
# round 2, call 17: bin-major mel stage of the STFT kernel: tests, then old (VIAI_STFT_MEL=walk) vs new timing
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_stft_gpu.py tests/test_librivox_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12 ) > gpurun_out/r02_pytest17.log 2>&1
tail -3 gpurun_out/r02_pytest17.log | cut -c1-300
( timeout 120 python scripts/r02_stft_time.py 2>&1 | tail -1 ) > gpurun_out/r02_stft17_new.log
( VIAI_STFT_MEL=walk timeout 120 python scripts/r02_stft_time.py 2>&1 | tail -1 ) > gpurun_out/r02_stft17_old.log
cat gpurun_out/r02_stft17_new.log gpurun_out/r02_stft17_old.log
