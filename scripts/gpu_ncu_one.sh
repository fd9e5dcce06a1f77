#!/bin/bash
# usage: gpu_ncu_one.sh <kernel-regex> <skip> <outname> [batch]
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/$3 -f \
    python bench.py --batch ${4:-2048} --steps 3 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$3.log 2>&1
ls -la gpurun_out/$3.ncu-rep
