"""Pipelined device-resident run with a per-stage timeline: python tools/timeline.py [stations] [steps] [mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import torch  # noqa: E402

import fmrx  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
iq = torch.randint(0, 256, (S, fmrx.BLOCK_BYTES), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
rx = fmrx.Batch(S, mode=mode, profile=fmrx.PROFILE_INTENT, max_blocks=1)
for _ in range(3):
    rx.process_device(iq.data_ptr(), 1, None)
rx.sync()
rx.profile(2)
for _ in range(steps):
    rx.process_device(iq.data_ptr(), 1, None)
for name, t0, t1 in rx.timeline():
    print(f"{name:14s} {t0:9.3f} {t1:9.3f}  ({t1 - t0:.3f})")
