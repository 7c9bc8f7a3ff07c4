#!/bin/bash
# 2 GPUs: overlapped data-parallel step
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/r02_ddp_overlap.py 2>&1 ) > gpurun_out/r02_ddp_overlap_n2.log 2>&1
grep -E "Error|error:|assert|OK|ms/step|File \"/tmp|line [0-9]+, in" gpurun_out/r02_ddp_overlap_n2.log | grep -v "site-packages/torch/distributed" | head -40 | cut -c1-300
