#!/bin/bash
# 2 GPUs: overlapped data-parallel step
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/r02_ddp_overlap.py 2>&1 | tail -30 ) > gpurun_out/r02_ddp_overlap_n2.log 2>&1
cat gpurun_out/r02_ddp_overlap_n2.log | cut -c1-300
