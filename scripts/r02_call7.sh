#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 ) > gpurun_out/r02_pytest7.log 2>&1
( timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench7.log 2>&1
( VIAI_FUSE_BWD_REDUCE=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-wavenet --no-extra 2>&1 | tail -2 ) > gpurun_out/r02_bench7_nofuse.log 2>&1
( timeout 300 python scripts/r02_op_table.py 2>&1 | tail -95 ) > gpurun_out/r02_op_table7.log 2>&1
( timeout 600 python scripts/c3_probe.py 2>&1 | tail -2 ) > gpurun_out/r02_c3_probe.log 2>&1
( timeout 600 python scripts/c3_kernels.py 2>&1 | tail -60 ) > gpurun_out/r02_c3_kernels.csv 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_mel_fast -s 1 -c 1 -o gpurun_out/r02_prof_stft -f python -c "
import sys; sys.path.insert(0,'.')
import torch
from viai_b200.utils import audio
y = torch.rand(16000*600, device='cuda')*2-1
for _ in range(3): audio.melspectrogram_cuda(y)
torch.cuda.synchronize()" > gpurun_out/r02_prof_stft.log 2>&1
tail -3 gpurun_out/r02_pytest7.log; for f in gpurun_out/r02_bench7.log gpurun_out/r02_bench7_nofuse.log; do grep -o '"ms_per_step": [0-9.]*' $f | head -1; done; head -3 gpurun_out/r02_op_table7.log | cut -c1-300; cat gpurun_out/r02_c3_probe.log | cut -c1-500; head -12 gpurun_out/r02_c3_kernels.csv
