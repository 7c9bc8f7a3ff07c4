#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 3 2>&1 ) > gpurun_out/r02_bench_n8_full.log 2>&1
grep "^{" gpurun_out/r02_bench_n8_full.log | cut -c1-300; grep -E "Error|error:|Traceback" gpurun_out/r02_bench_n8_full.log | head -5; tail -4 gpurun_out/r02_bench_n8_full.log | grep real
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 scripts/r02_ddp_overlap.py 2>&1 ) > gpurun_out/r02_ddp_overlap_n8.log 2>&1
grep -E "OK|ms/step|Error" gpurun_out/r02_ddp_overlap_n8.log | head
