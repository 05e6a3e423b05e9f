"""Small device-resident run of the chain for ncu: python tools/prof_chain.py [stations] [steps] [reference|strict|fma] [synth|random]

Input: the bench's synthetic FM multiplex by default (loops lock, the PLL kernel runs its fast path as in the bench -- profile a step
after the first, `ncu -s`); `random` = uniform random bytes (every loop chases noise: the libm redo path shows up)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import torch  # noqa: E402

import fmrx  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if (sys.argv[4] if len(sys.argv) > 4 else "synth") == "random":
    iq = torch.randint(0, 256, (S, fmrx.BLOCK_BYTES), dtype=torch.uint8, device=dev)
else:
    from fmrx import synth  # noqa: E402
    iq = synth.synth_batch_torch(range(S), 1, 0, dev, chunk=64)
torch.cuda.synchronize()
num = {"reference": fmrx.NUMERICS_REFERENCE, "strict": fmrx.NUMERICS_STRICT, "fma": fmrx.NUMERICS_FMA}[sys.argv[3] if len(sys.argv) > 3 else "reference"]
rx = fmrx.Batch(S, mode=0, profile=fmrx.PROFILE_INTENT, max_blocks=1, numerics=num)
for _ in range(steps):
    rx.process_device(iq.data_ptr(), 1, None)
rx.sync()
print("ok", rx.launches)
