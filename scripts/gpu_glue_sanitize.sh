#!/bin/bash
# glue tests (incl. controller-thread selection) + compute-sanitizer over the stage tests and the closed-loop smoke
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_glue.py -q -s -m gpu 2>&1 | tail -12 > gpurun_out/pytest_glue.log; cat gpurun_out/pytest_glue.log
bash scripts/gpu_sanitize.sh
python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-other-configs --cfo-hz 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('clocks', d['clocks'], 'value', d['value'])"
