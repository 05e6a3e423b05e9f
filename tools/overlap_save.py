"""One measured comparison: the 151-tap D = 1 filters as overlap-save FFT convolution
(cuFFT through torch.fft -- a LIBRARY path, measured for the comparison only, never adopted into the chain) against the
direct register-tiled kernel (`fir151_kernel<1, ...>`: 0.305 ms fused-multiply-add, 0.585 ms reference-exact, 4096 stations x
15360 samples).  Reports time per filter pass and the error against a float64 direct convolution.

    python tools/overlap_save.py [stations] > gpurun_out/<tag>_overlap_save.txt
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import fmrx  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N, T = 15360, 151
dev = torch.device("cuda", 0)
h = torch.tensor(fmrx.design_bpf(22e3, 54e3, 240e3, T), device=dev)
x = torch.randn(S, N, device=dev) * 0.3
hist = torch.zeros(S, T - 1, device=dev)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, y


def overlap_save(seg):
    """segments of `seg` outputs, FFT length seg + 150 rounded up to a power of two friendly size"""
    L = seg + T - 1
    nfft = 1 << (L - 1).bit_length()
    H = torch.fft.rfft(h, nfft)
    xe = torch.cat([hist, x], 1)                                   # [S][N + 150]
    nseg = N // seg
    idx = (torch.arange(nseg, device=dev) * seg)[:, None] + torch.arange(L, device=dev)[None, :]

    def run():
        blocks = xe[:, idx]                                         # [S][nseg][L] (gather: the overlap is read twice)
        Y = torch.fft.irfft(torch.fft.rfft(blocks, nfft) * H, nfft)
        return Y[:, :, T - 1:L].reshape(S, N)
    return run, nfft


ref = None
if S <= 64:
    xe = torch.cat([hist, x], 1).double().cpu().numpy()
    ref = np.stack([np.convolve(xe[s], h.double().cpu().numpy())[T - 1:T - 1 + N] for s in range(S)])
print(f"overlap-save vs direct, {S} stations x {N} samples, {T} taps (one D = 1 filter pass); direct kernel: 0.305 ms FFMA / 0.585 ms reference-exact at 4096 stations")
for seg_eff in (512, 1920, 3840, 7680, 15360):  # divisors of 15360; FFT lengths 1024, 4096, 4096, 8192, 16384
    run, nfft = overlap_save(seg_eff)
    try:
        ms, y = timed(run)
    except RuntimeError as e:  # out of memory at the largest sizes
        print(f"  segment {seg_eff:6d} (FFT {nfft:6d}): {str(e)[:60]}")
        continue
    err = ""
    if ref is not None:
        d = y.double().cpu().numpy() - ref
        err = f"  rel-rms vs float64 direct {np.sqrt(np.mean(d ** 2)) / np.sqrt(np.mean(ref ** 2)):.3g}"
    print(f"  segment {seg_eff:6d} (FFT {nfft:6d}): {ms:8.3f} ms{err}")
# the direct kernel on the same data, through the C-ABI operator (host buffers: timed without the copies by the bench's stage pass;
# here only its error is of interest)
if ref is not None:
    zi = np.zeros((S, 150), np.float32)
    yd = fmrx.fir_decim(x.cpu().numpy().reshape(S, 1, N), h.cpu().numpy(), zi, 1, exact=False).reshape(S, N)
    d = yd.astype(np.float64) - ref
    print(f"  direct FFMA kernel: rel-rms vs float64 direct {np.sqrt(np.mean(d ** 2)) / np.sqrt(np.mean(ref ** 2)):.3g}")
