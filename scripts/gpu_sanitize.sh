#!/bin/bash
# compute-sanitizer over the stage tests and a short closed-loop run -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k "ofdm_demod_bit_exact or find_index or viterbi_bit_exact or fic_decode or rs_superframes" > gpurun_out/sanitizer_${tool}_stages.log 2>&1
  tail -4 gpurun_out/sanitizer_${tool}_stages.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  tail -4 gpurun_out/sanitizer_${tool}_smoke.log
done
