#!/bin/bash
# pipelined-path parity test, then the default bench's e2e part on a reduced batch (quick), then the full default bench
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_r2.py -q -m gpu -k "pipelined" 2>&1 | tail -6
python bench.py --no-cpu-baseline --no-other-configs --cfo-hz 0 --e2e-batch 2048 2>gpurun_out/bench_e2e_small.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('value', round(d['value']), 'clocks', d['clocks']); print({k: e[k] for k in e if k not in ('cf32_input','u8_input','note')}); print('u8', {k: e['u8_input'][k] for k in e['u8_input'] if k != 'note'})"
tail -3 gpurun_out/bench_e2e_small.err
