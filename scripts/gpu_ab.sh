#!/bin/bash
# A/B of an environment switch, alternating runs (the first bench of a session runs a few % slower than the following ones):
#   gpu_ab.sh VAR=value [repeats]
mkdir -p gpurun_out
one() {
    python bench.py --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 --steps 20 > gpurun_out/bench_ab_$1.json 2> gpurun_out/bench_ab_$1.err
    python - "$1" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_ab_{sys.argv[1]}.json").read().strip().splitlines()[-1]); k = d["kernels"]
print(sys.argv[1], "|", round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms/step | ofdm", round(k["ofdm_demod_kernel"]["ms_per_step"], 3), "vit MSC", round(k["viterbi_kernel(MSC)"]["ms_per_step"], 3), flush=True)
PY
}
one warmup
for i in $(seq 1 ${2:-3}); do
    one base$i
    env "$1" bash -c "$(declare -f one); one alt$i"
done
