#!/bin/bash
# device-resident bench under a list of environment settings, one line each:
#   bash tools/gpu_sweep.sh <tag> "FMRX_PLL_STEP=0" "FMRX_PLL_STEP=2 FMRX_PLL_SMS=24" ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for envs in "$@"; do
  echo "== $envs"
  env $envs python bench.py --device-only --no-check --steps ${SWEEP_STEPS:-40} 2>&1 | tail -1
done 2>&1 | tee $OUT/${TAG}_sweep.txt
