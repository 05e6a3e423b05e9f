"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/memcheck_chain.py
Ragged batch (33 stations), two calls of two blocks, every mode x numerics x back end, the quality profile, the ring, and the
function-level operators with sizes that leave partial tiles."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import fmrx  # noqa: E402

F = np.float32
rng = np.random.default_rng(0)
S, B = 33, 2
for mode in (0, 1, 2):
    raw = rng.integers(0, 256, (S, 2 * B * fmrx.BLOCK_BYTES), dtype=np.uint8)
    for numerics in (fmrx.NUMERICS_REFERENCE, fmrx.NUMERICS_STRICT, fmrx.NUMERICS_FMA):
        for extra in (0, fmrx.PATH_RDS_STAGES):
            for quality in (0, 13):
                if mode == 1 and extra:
                    continue
                with fmrx.Batch(S, mode=mode, profile=1, max_blocks=B, paths=fmrx.PATH_AUDIO | fmrx.PATH_RDS | extra, numerics=numerics, quality=quality) as rx:
                    rx.process(raw[:, :B * fmrx.BLOCK_BYTES], want_float=True)
                    rx.process(raw[:, B * fmrx.BLOCK_BYTES:])
                    blob = rx.get_state()
                    rx.reset()
                    rx.set_state(blob)
    print("mode", mode, "ok", flush=True)
with fmrx.Batch(5, mode=0, profile=0, max_blocks=1) as rx, fmrx.Ring(rx, n_slots=2, n_blocks=1) as ring:
    for k in range(3):
        ring.acquire()[:] = rng.integers(0, 256, (5, fmrx.BLOCK_BYTES), dtype=np.uint8)
        ring.commit()
        assert ring.next() is not None
        ring.release()
    ring.close()
h = fmrx.design_bpf(113.5e3, 114.5e3, 240e3, 151)
for n in (152, 1000, 2500):
    x = rng.standard_normal((3, 2, n)).astype(F)
    fmrx.pll_combine(x, h, np.zeros((3, 150), F), 114000, 240000, 0.5, 0.1, 0.001, np.tile(np.array([0, 0, 1, 0, 0, 1], F), (3, 1)))
    fmrx.fir_decim(x, h, np.zeros((3, 150), F), 1, exact=True)
    fmrx.fir_mixer(x, x, h, np.zeros((3, 150), F))
fmrx.deemphasis(rng.standard_normal((3, 2, 2 * 37)).astype(F), 75.0, 48000.0, np.zeros((3, 4), F))
st = np.zeros((40, fmrx.RDS_STATE_WORDS), np.int32)
fmrx.rds_decode(rng.standard_normal((40, 2, 3648)).astype(F), st)
print("all ok")
