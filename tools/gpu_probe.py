"""Prints device facts and the FP32 issue-rate microbenchmarks (roofline denominators for the FIR kernels)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import fmrx  # noqa: E402

names = {0: "FFMA", 1: "FMUL+FADD", 2: "FFMA2 (packed)", 3: "FMUL2+FADD2 (packed)"}
out = {names[k]: round(fmrx.measure_fp32_peak(k, reps=5), 3) for k in range(4)}
print(json.dumps({"fp32_tera_lane_ops_per_s": out}))
