#!/bin/bash
# final library on 2 GPUs: the bench line (weak + strong scaling sub-record), NCCL all-reduces overlapped inside the captured step
mkdir -p gpurun_out
( time timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-wavenet 2>&1 ) > gpurun_out/r02_bench_n2_final.log 2>&1
grep "^{" gpurun_out/r02_bench_n2_final.log | cut -c1-400; grep real gpurun_out/r02_bench_n2_final.log; grep -i "error\|Traceback" gpurun_out/r02_bench_n2_final.log | head -5
