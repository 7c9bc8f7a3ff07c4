#!/bin/bash
mkdir -p gpurun_out
( for f in 4 12; do echo "=== flags $f ==="; timeout 200 python scripts/tc_probe.py $f 2>&1 | tail -20; done ) > gpurun_out/probe3.log 2>&1
cat gpurun_out/probe3.log
