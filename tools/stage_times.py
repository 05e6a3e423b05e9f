"""Per-stage device times of one mode (phases serialised on one stream): python tools/stage_times.py [mode] [stations] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import torch  # noqa: E402

import fmrx  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
iq = torch.randint(0, 256, (S, fmrx.BLOCK_BYTES), dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
rx = fmrx.Batch(S, mode=mode, profile=fmrx.PROFILE_INTENT, max_blocks=1)
for _ in range(3):
    rx.process_device(iq.data_ptr(), 1, None)
rx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s_first = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream_phase(rx.h, 0), device=dev)
s_last = torch.cuda.ExternalStream(fmrx.lib().fmrx_batch_cuda_stream(rx.h), device=dev)
e0.record(s_first)
for _ in range(steps):
    rx.process_device(iq.data_ptr(), 1, None)
e1.record(s_last)
rx.sync()
print(f"mode {mode}: pipelined {e0.elapsed_time(e1) / steps:.3f} ms per step, partition {rx.partition()}")
rx.profile(True)
for _ in range(steps):
    rx.process_device(iq.data_ptr(), 1, None)
st = rx.stage_times()
rx.profile(False)
tot = 0.0
for k, (ms, cnt) in st.items():
    if cnt:
        print(f"  {k:14s} {ms / steps:8.4f} ms")
        tot += ms / steps
print(f"  {'sum':14s} {tot:8.4f} ms")
