#!/bin/bash
# GPU call: parity tests, bench lines, extras (WaveNet / STFT), ncu launch list + one full capture of the tensor-core conv
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 400 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench1.log 2>&1
( time timeout 200 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
( time timeout 300 python scripts/bench_extra.py 16000 ) > gpurun_out/bench_extra.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01_tc.csv \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 20 -c 6 -o gpurun_out/prof_conv_tc_r01 \
   python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench1.log; tail -2 gpurun_out/bench_extra.log
