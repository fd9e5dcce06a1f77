#!/bin/bash
# Builds experimental variants of libdab_b200.so (compile-time switches) next to the repo root as gpurun_exp_<name>.so (git-ignored; they
# travel to the GPU box with the snapshot).  Run on the CPU box, then on the GPU:   bash scripts/gpu_variants.sh
#   pairs     -DDABB_DEMAP_PAIRS           two carriers per packed reciprocal chain in the demap   (round 1: -1 % on ofdm_demod_kernel)
#   cta6      -DDEMOD_CTAS_PER_SM=6        6 CTAs/SM: 80 registers (spills), 37 KB shared memory (twiddles through L1, softbit staging inside
#                                          the exchange buffer at the price of two more barriers per symbol)
#   cta6pairs both
#   vit64     -DVIT_THREADS_N=64           Viterbi CTAs of 64 codewords                              (round 1: no change)
set -eu
cd "$(dirname "$0")/../welle.io_b200/csrc"
make -s
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-ffp-contract=off"
build() {   # name, source, macro
    nvcc $FLAGS $3 -c $2.cu -o build/$2_$1.o
    objs=""
    for o in api ofdm viterbi rs tables; do if [ $o = $2 ]; then objs="$objs build/$2_$1.o"; else objs="$objs build/$o.o"; fi; done
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../gpurun_exp_$1.so $objs -lcudart
    echo "built gpurun_exp_$1.so"
}
build nopairs ofdm -DDABB_NO_DEMAP_PAIRS
build vittfma viterbi -DVIT_T_FMA
