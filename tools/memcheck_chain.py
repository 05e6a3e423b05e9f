"""Small runs of every kernel family in every configuration: ragged batch (33 stations), two calls of two blocks, every mode x
numerics x RDS back end x quality profile with a checkpoint / resume in between, the ring, and the function-level operators with
sizes that leave partial tiles.  Written as a driver for compute-sanitizer (`compute-sanitizer --tool memcheck python
tools/memcheck_chain.py`); on pools where the sanitizer is not available it still runs as a configuration sweep
(tests/test_gpu_configs.py) and checks what must hold across configurations: int16 audio is identical for REFERENCE and STRICT
numerics and for both RDS back ends, and the RDS bits of both back ends agree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))
import fmrx  # noqa: E402

F = np.float32
rng = np.random.default_rng(0)
S, B = 33, 2
sys.path.insert(0, ROOT)
from fmrx import synth  # noqa: E402

for mode in (0, 1, 2):
    raw = np.stack([synth.synth_station(s % 4, 2 * B, mode) for s in range(S)])
    seen = {}
    for numerics in (fmrx.NUMERICS_REFERENCE, fmrx.NUMERICS_STRICT, fmrx.NUMERICS_FMA):
        for extra in (0, fmrx.PATH_RDS_STAGES):
            for quality in (0, 13):
                if mode == 1 and extra:
                    continue
                with fmrx.Batch(S, mode=mode, profile=1, max_blocks=B, paths=fmrx.PATH_AUDIO | fmrx.PATH_RDS | extra, numerics=numerics, quality=quality) as rx:
                    r1 = rx.process(raw[:, :B * fmrx.BLOCK_BYTES], want_float=True)
                    blob = rx.get_state()
                    rx.reset()
                    rx.set_state(blob)
                    r2 = rx.process(raw[:, B * fmrx.BLOCK_BYTES:])
                    audio = np.concatenate([r1["audio"], r2["audio"]], 1)
                    assert np.isfinite(r1["audio_f"][~np.isnan(r1["audio_f"])]).all()
                    if numerics != fmrx.NUMERICS_FMA:
                        ref = seen.setdefault(("audio", quality), audio)
                        assert np.array_equal(ref, audio), f"mode {mode} numerics {numerics} paths+{extra} quality {quality}: int16 audio differs between configurations"
                    if mode != 1:
                        bits = np.concatenate([r1["rds_bits"], r2["rds_bits"]], 1)
                        refb = seen.setdefault(("bits", quality), bits)
                        assert np.array_equal(refb, bits), f"mode {mode} numerics {numerics} paths+{extra} quality {quality}: RDS bits differ between configurations"
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
