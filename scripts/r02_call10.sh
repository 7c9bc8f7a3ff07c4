#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 ) > gpurun_out/r02_bench_n2_full.log 2>&1
grep "^{" gpurun_out/r02_bench_n2_full.log | cut -c1-400; grep -E "Error|error:|Traceback" gpurun_out/r02_bench_n2_full.log | head -5; tail -4 gpurun_out/r02_bench_n2_full.log | grep real
