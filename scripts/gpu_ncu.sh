#!/bin/bash
# ncu captures for profiles/: launch list of one bench run + full sets of the two hot kernels
set -u
mkdir -p gpurun_out
B=${1:-2048}
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:dabb -s 70 -c 70 --csv --log-file gpurun_out/launches.csv \
    python bench.py --batch $B --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --cfo-hz 0 > gpurun_out/ncu_bench1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ofdm_demod_kernel -s 9 -c 1 -o gpurun_out/prof_ofdm -f \
    python bench.py --batch $B --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --cfo-hz 0 > gpurun_out/ncu_bench2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:viterbi_kernel -s 19 -c 1 -o gpurun_out/prof_viterbi -f \
    python bench.py --batch $B --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --cfo-hz 0 > gpurun_out/ncu_bench3.log 2>&1
# the same kernel with the oscillator active (carrier offset run of bench.py: launches 21.. of the regex)
ncu --set full --clock-control none --import-source on -k regex:ofdm_demod_kernel -s 28 -c 1 -o gpurun_out/prof_ofdm_nco -f \
    python bench.py --batch $B --steps 3 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench4.log 2>&1
ls -la gpurun_out/
