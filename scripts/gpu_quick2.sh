#!/bin/bash
mkdir -p gpurun_out
one() {
    python bench.py --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 --steps 10 > gpurun_out/bench_var_$1.json 2> gpurun_out/bench_var_$1.err
    python - "$1" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_var_{lib}.json").read().strip().splitlines()[-1]); k = d["kernels"]
    print(lib, "|", round(d["value"]), "frames/s", round(d["ms_per_step"], 3), "ms/step | ofdm", round(d["roofline"]["ms_per_launch"], 4), "| viterbi FIC", round(k["viterbi_kernel(FIC)"]["ms_per_step"], 3), "MSC", round(k["viterbi_kernel(MSC)"]["ms_per_step"], 3), flush=True)
except Exception as e:
    print(lib, "| bench failed:", e, flush=True)
PY
}
DABB_LANES=partition python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -3
one burst
DABB_LANES=partition one partition
DABB_LANES=partition DABB_TRACE=$PWD/gpurun_out/trace_part.txt python bench.py --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 --steps 4 > /dev/null 2>&1
