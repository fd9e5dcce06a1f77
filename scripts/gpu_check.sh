#!/bin/bash
# one GPU-box session: parity tests, smoke, bench (small + full), reference arm, ncu captures -> gpurun_out/
set -u
mkdir -p gpurun_out
B=${1:-8192}
echo "== pytest -m gpu"; python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench small"; python bench.py --batch 512 --steps 3 --warmup 6 --no-cpu-baseline --e2e-batch 128 2>gpurun_out/bench_small.err | tee gpurun_out/bench_small.json | cut -c1-1500
tail -5 gpurun_out/bench_small.err
echo "== bench full"; python bench.py --batch $B --steps 10 --warmup 8 2>gpurun_out/bench_full.err | tee gpurun_out/bench_full.json | cut -c1-6000
tail -5 gpurun_out/bench_full.err
