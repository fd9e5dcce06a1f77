#!/bin/bash
# ncu captures for profiles/: launch list of one bench run + full sets (with source) of the hot kernels, batch $1 (default 8192)
set -u
mkdir -p gpurun_out
B=${1:-8192}
BENCH="python bench.py --batch $B --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --no-other-configs"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'dabb|StreamState|StepScratch' -s 70 -c 56 --csv --log-file gpurun_out/launches.csv \
    $BENCH --cfo-hz 0 > gpurun_out/ncu_bench1.log 2>&1
# steady-state launches: step 9 (0-based) of the bench = 10th ofdm launch; MSC Viterbi = 20th viterbi launch
ncu --set full --clock-control none --import-source on -k regex:ofdm_demod_kernel -s 9 -c 1 -o gpurun_out/prof_ofdm -f $BENCH --cfo-hz 0 > gpurun_out/ncu_bench2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 19 -c 1 -o gpurun_out/prof_viterbi -f $BENCH --cfo-hz 0 > gpurun_out/ncu_bench3.log 2>&1
ls -la gpurun_out/*.ncu-rep
