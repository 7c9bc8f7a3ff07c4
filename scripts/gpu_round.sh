#!/bin/bash
# Round-end style pass on one B200: GPU tests, smoke(), both bench arms, ncu launch list + one full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_n1.log 2>&1
( time timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none -k "regex:conv_tc_kernel" --launch-skip 22 --launch-count 1 -o gpurun_out/prof_one -f \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-wavenet > gpurun_out/ncu_one.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -1 gpurun_out/bench_n1.log | cut -c1-400
