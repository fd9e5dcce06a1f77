#!/bin/bash
# ncu --set full of side kernels of one steady-state step; usage: gpu_ncu_side.sh kernel_regex ...
set -u
mkdir -p gpurun_out
BENCH="python bench.py --batch 8192 --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0"
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -o gpurun_out/prof_$k -f $BENCH > gpurun_out/ncu_side_$k.log 2>&1
done
ls -la gpurun_out/prof_*kernel.ncu-rep
