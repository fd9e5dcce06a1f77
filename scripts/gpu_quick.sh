#!/bin/bash
# quick GPU check: stage + pipeline + glue parity tests, then the short variant bench of the default library and gpurun_exp_*.so
mkdir -p gpurun_out
if [ -z "${SKIP_TESTS:-}" ]; then python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_glue.py tests/test_gpu_fullsize.py -m gpu -q -s 2>&1 | tail -25 > gpurun_out/pytest_quick.log; tail -3 gpurun_out/pytest_quick.log; fi
one() {
    python bench.py --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 --steps 10 > gpurun_out/bench_var_$1.json 2> gpurun_out/bench_var_$1.err
    python - "$1" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_var_{lib}.json").read().strip().splitlines()[-1]); k = d["kernels"]
    print(lib, "|", round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms/step | ofdm", round(d["roofline"]["ms_per_launch"], 4), "ms frac", round(d["roofline"]["frac"], 4),
          "survey-frac", round(d["roofline"]["frac_with_survey_bytes"], 4), "| viterbi FIC", round(k["viterbi_kernel(FIC)"]["ms_per_step"], 3), "MSC", round(k["viterbi_kernel(MSC)"]["ms_per_step"], 3),
          "gather", round(k["msc_gather_kernel"]["ms_per_step"], 3), flush=True)
except Exception as e:
    print(lib, "| bench failed:", e, flush=True)
PY
}
unset DABB_LIB
one default
DABB_CORESIDENT=1 DABB_CORESIDENT_SERIAL=1 one coresident_serial
for lib in $(ls gpurun_exp_*.so 2>/dev/null); do export DABB_LIB=$PWD/$lib; one "$lib"; done
