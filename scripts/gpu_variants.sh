#!/bin/bash
# On the GPU box: the stage + pipeline parity tests and a short bench for every gpurun_exp_*.so built by scripts/build_variants.sh
# (and for the default library first).  One line per variant: frames/s, ms/step, ofdm_demod_kernel ms and roofline fraction.
mkdir -p gpurun_out
for lib in default $(ls gpurun_exp_*.so 2>/dev/null); do
    if [ "$lib" = default ]; then unset DABB_LIB; else export DABB_LIB=$PWD/$lib; fi
    t=$(python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -q 2>&1 | tail -1)
    python bench.py --no-cpu-baseline --no-e2e --cfo-hz 50 > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
    python - "$lib" "$t" <<'PY'
import json, sys
lib, t = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/bench_{lib}.json").read().strip().splitlines()[-1])
    o = d["roofline"].get("oscillator_active") or {}
    print(lib, "|", t, "|", round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms/step, ofdm", round(d["roofline"]["ms_per_launch"], 3), "ms frac", round(d["roofline"]["frac"], 3),
          "| oscillator active: ofdm", round(o.get("ofdm_ms_per_launch", 0), 3), "ms")
except Exception as e:
    print(lib, "|", t, "| bench failed:", e)
PY
done
