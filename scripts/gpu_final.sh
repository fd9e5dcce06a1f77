#!/bin/bash
# round-end sequence + refreshed ncu captures of the hot kernels, one GPU-box session
set -u
bash scripts/gpu_check.sh
bash scripts/gpu_r2_profiles.sh > gpurun_out/r2_profiles.log 2>&1
tail -n 8 gpurun_out/r2_profiles.log
