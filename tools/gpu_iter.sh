#!/bin/bash
# quick GPU iteration: parity suite + device-only bench under a few PLL partition sizes.  bash tools/gpu_iter.sh <tag> [sms...]
TAG=${1:-it}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
for sms in "${@:-32}"; do
  FMRX_PLL_SMS=$sms python bench.py --device-only --no-check 2>&1 | tee -a $OUT/${TAG}_sweep.log
done
