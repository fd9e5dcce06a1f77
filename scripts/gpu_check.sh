#!/bin/bash
# one GPU-box session, the round-end sequence: parity tests, smoke, default bench, reference arm, launch list -> gpurun_out/
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; SECONDS=0; python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full.log 2>&1; tail -4 gpurun_out/pytest_full.log; echo "pytest: ${SECONDS}s"
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"; SECONDS=0; python bench.py 2>gpurun_out/bench_full.err > gpurun_out/bench_full.json; echo "bench: ${SECONDS}s rc=$?"; cut -c1-600 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== reference arm"; SECONDS=0; python bench.py --impl reference 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json; echo "reference arm: ${SECONDS}s"; cut -c1-900 gpurun_out/bench_ref.json
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'dabb|StreamState|StepScratch' -s 70 -c 56 --csv --log-file gpurun_out/launches.csv \
    python bench.py --batch 8192 --steps 3 --warmup 8 --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 > gpurun_out/ncu_bench1.log 2>&1
grep -c . gpurun_out/launches.csv
