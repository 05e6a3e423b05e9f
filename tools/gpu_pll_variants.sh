#!/bin/bash
# A/B of the PLL step variants (FMRX_PLL_STEP, csrc/fmrx_pll.cu) on one GPU: pipe microbenchmarks (do the FP64 and the
# conversion pipe overlap?), chain latency per variant, parity of the chain tests per variant, device-resident bench per
# variant and partition size.
#   bash tools/gpu_pll_variants.sh <tag>
TAG=${1:-pllv}
OUT=gpurun_out; mkdir -p $OUT
python - <<'PY' 2>&1 | tee $OUT/${TAG}_micro.txt
import os, subprocess, sys, json
sys.path.insert(0, "real-time-software-defined-radio_b200")
import fmrx
r = {k: fmrx.measure_fp32_peak(k) for k in (0, 4, 5, 6, 7)}
print("tera lane-ops/s: ffma %.2f dfma %.2f f2f %.2f alu %.2f mixed(4 dfma + 2 f2f) %.2f" % (r[0], r[4], r[5], r[6], r[7]))
t_serial = 4 / r[4] + 2 / r[5]; t_overlap = max(4 / r[4], 2 / r[5])
print("mixed: measured %.3f per unit; serialised pipes would give %.3f, overlapped %.3f" % (6 / r[7], t_serial, t_overlap))
for v in (0, 1, 2):
    out = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, 'real-time-software-defined-radio_b200'); import fmrx; print(fmrx.measure_pll_chain())"],
                         env=dict(os.environ, FMRX_PLL_STEP=str(v)), capture_output=True, text=True)
    print("chain latency, variant", v, out.stdout.strip(), out.stderr.strip()[-200:])
PY
for v in 1 2; do
  FMRX_PLL_STEP=$v python -m pytest tests/test_gpu_chain.py tests/test_gpu_functions.py -m gpu -x -q -k "golden or noise or pll or sweep" > $OUT/${TAG}_tests_v$v.log 2>&1; echo "variant $v tests rc=$?"; tail -2 $OUT/${TAG}_tests_v$v.log
done
for v in 0 1 2; do
  for sms in "" 24; do
    echo "== FMRX_PLL_STEP=$v FMRX_PLL_SMS=${sms:-default}"
    if [ -z "$sms" ]; then FMRX_PLL_STEP=$v python bench.py --device-only --no-check --steps 40 2>&1 | tail -1
    else FMRX_PLL_STEP=$v FMRX_PLL_SMS=$sms python bench.py --device-only --no-check --steps 40 2>&1 | tail -1; fi
  done
done 2>&1 | tee $OUT/${TAG}_bench.txt
