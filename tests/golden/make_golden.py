"""Generates tests/golden/*.npz by running the UNMODIFIED reference here (oracle/_ref: the reference binary and the
reference objects behind oracle/ref_shim.cpp).  Run once in the authoring container (where /root/reference is mounted):

    make -C oracle && python tests/golden/make_golden.py

The fixtures are what pins the oracle port (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_*.py) on
machines where /root/reference does not exist.  Inputs are re-synthesised from seeds by fmrx.synth at test time; each
fixture stores the SHA-256 of the input it was generated from so generator drift is detected, not silently absorbed.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))

from fmrx import synth  # noqa: E402
from oracle import Ref, RefChain, run_ref_binary  # noqa: E402

F = np.float32
LONG_STRIDE = 4  # 15360-sample taps are stored every 4th sample


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_text(stderr_text, nblk):
    """The frame_thread part of the binary's stderr, cut after `nblk` whole blocks (the binary runs one extra block
    on the short read at EOF)."""
    lines = stderr_text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("initial offset"))
    out = []
    for l in lines[start:]:
        if l.startswith("****************Prcoessing Block: %d*" % nblk) or l.startswith("Run: gnuplot"):
            if out and out[-1] == " ":
                out.pop()
            break
        out.append(l)
    return "\n".join(out) + "\n"


def chain_fixture(mode, nblk, seed, long_blocks):
    raw = synth.synth_iq(nblk, mode, seed=seed)
    audio_bin, err, rc = run_ref_binary(raw, mode)
    assert rc == 0
    d = dict(mode=mode, nblk=nblk, seed=seed, input_sha256=sha(raw))
    npb = 2 * (3072 if mode == 0 else 2949)
    d["binary_audio"] = audio_bin[:nblk * npb]
    if mode == 0:
        d["binary_frame_text"] = frame_text(err, nblk)
    for profile, name in ((0, "binary"), (1, "intent")):
        ch = RefChain(mode, profile)
        audio, rrcs = [], []
        for b in range(nblk):
            audio.append(ch.block(raw[b * 307200:(b + 1) * 307200]))
            for t in ("mono", "stereo", "audio_f") + (("rds_rrc",) if mode == 0 else ()):
                d[f"{name}_{t}_{b}"] = ch.taps[t]
            if b in long_blocks and (profile == 1 or b == 0):
                for t in ("demod", "pilot", "nco", "stereo_bpf") + (("rds_bpf", "rds_sq", "rds_nco", "rds_lpf", "rds_res") if mode == 0 else ()):
                    v = ch.taps[t]
                    d[f"{name}_{t}_{b}"] = v[::LONG_STRIDE] if v.size >= 15360 else v
            if mode == 0:
                rrcs.append(ch.taps["rds_rrc"])
        d[f"{name}_audio"] = np.concatenate(audio)
        if profile == 0:
            assert np.array_equal(d["binary_audio"], np.concatenate(audio)), "RefChain(binary) != fm_radio stdout"
        if mode == 0 and profile == 1:
            d["frame_text"] = Ref().frame_thread(np.stack(rrcs))
            assert d["frame_text"] == d["binary_frame_text"], "frame_thread(shim) != fm_radio stderr"
    return d


def func_fixture():
    r = Ref()
    rng = np.random.default_rng(20261018)
    d = {}
    # --- designs used by src/fm_radio.cpp (:40-42,200-203,366-370) plus the function-level 44.1 kHz case
    d["lpf_rf0"] = r.lpf(2.4e6, 1e5, 151); d["lpf_rf1"] = r.lpf(2.5e6, 1e5, 151)
    d["lpf_mono0"] = r.lpf(240000, 16000, 151); d["lpf_mono1"] = r.lpf(6e6, 16000, 3624)
    d["lpf_3k"] = r.lpf(240000, 3000, 151); d["lpf_anti"] = r.lpf(float(F(240000) * F(19)), 28500, 2869)
    d["lpf_441"] = r.lpf(240000 * 147, 16000, 151 * 147)
    d["bpf_pilot0"] = r.bpf(18.5e3, 19.5e3, 240000, 151); d["bpf_stereo0"] = r.bpf(22e3, 54e3, 240000, 151)
    d["bpf_pilot1"] = r.bpf(18.5e3, 19.5e3, 6e6, 151); d["bpf_stereo1"] = r.bpf(22e3, 54e3, 6e6, 151)
    d["bpf_rds"] = r.bpf(54000, 60000, 240000, 151); d["bpf_sq"] = r.bpf(113500, 114500, 240000, 151)
    d["rrc"] = r.rrc(57000, 151)
    # --- unpack: all 256 byte values, and a short read (Q9)
    d["unpack_in"] = np.arange(256, dtype=np.uint8); d["unpack_out"] = r.unpack(d["unpack_in"])
    d["unpack_short_out"] = r.unpack(d["unpack_in"][:100], 256)
    # --- stateful FIR, three consecutive blocks each, state carried
    nb, n = 3, 2400
    x = rng.standard_normal((nb, n)).astype(F); xq = rng.standard_normal((nb, n)).astype(F)
    d["fir_x"], d["fir_xq"] = x, xq
    for decim in (1, 5, 10):
        zi = np.zeros(150, F)
        d[f"fir_d{decim}"] = np.stack([r.fir_decim(x[b], d["lpf_mono0"], zi, decim) for b in range(nb)])
        zi = np.zeros(150, F)
        assert np.array_equal(d[f"fir_d{decim}"], np.stack([r.fir_decim_ptr(x[b], d["lpf_mono0"], zi, decim) for b in range(nb)]))
    zi, zq = np.zeros(150, F), np.zeros(150, F)
    iq = [r.fir_decim_iq(x[b], xq[b], d["lpf_rf0"], zi, zq, 10) for b in range(nb)]
    d["fir_iq_i"] = np.stack([a for a, _ in iq]); d["fir_iq_q"] = np.stack([b for _, b in iq])
    d["demod"] = np.stack([r.demod(d["fir_iq_i"][b], d["fir_iq_q"][b]) for b in range(nb)])
    # --- resamplers: mode-1 mono 24/125 (NaN tap, Q5), RDS 19/80 (x19), 44.1 kHz 147/800, mode-1 stereo (5,24) truncated
    xr = rng.standard_normal((nb, 15360)).astype(F)
    d["res_x"] = xr
    zi = np.zeros(3623, F); d["res_24_125"] = np.stack([r.resample_ptr(xr[b], d["lpf_mono1"], zi, 125, 24) for b in range(nb)])
    zi = np.zeros(2868, F); d["res_19_80"] = np.stack([r.resample_rds(xr[b], d["lpf_anti"], zi, 80, 19) for b in range(nb)])
    # 44.1 kHz (BASELINE config 2) is reachable at function level only; with the reference's zi = taps-1 = 22196 convention
    # the state update needs blocks longer than that, so the case uses 100 ms blocks: 24000 in -> 4410 out
    x441 = rng.standard_normal((nb, 24000)).astype(F); d["res_x441"] = x441
    zi = np.zeros(151 * 147 - 1, F); d["res_147_800"] = np.stack([r.resample_ptr(x441[b], d["lpf_441"], zi, 800, 147) for b in range(nb)])
    zi = np.zeros(3623, F); d["res_24_5"] = np.stack([r.resample(xr[b], d["lpf_mono1"], zi, 5, 24, ny_keep=2949) for b in range(nb)])
    # --- PLLs on noisy carriers
    k = np.arange(nb * 6000)
    pil = (0.3 * np.cos(2 * np.pi * 19000 / 240000 * k + 0.7) + 0.01 * rng.standard_normal(k.size)).astype(F).reshape(nb, -1)
    d["pll_x"] = pil
    st = np.array([0, 0, 1, 0, 0, 1], F)
    d["pll_nco"] = np.stack([r.pll(pil[b], 19e3, 240e3, 2.0, 0.0, 0.01, st) for b in range(nb)]); d["pll_state"] = st.copy()
    sub = (0.2 * np.cos(2 * np.pi * 57000 / 240000 * k + 0.3) * np.sign(np.sin(2 * np.pi * k / 2000.0)) + 0.01 * rng.standard_normal(k.size)).astype(F).reshape(nb, -1)
    d["pllc_x"] = sub
    st = np.array([0, 0, 1, 0, 0, 1], F); zi = np.zeros(150, F); zl = np.zeros(150, F)
    ph = float(F(float(F(np.pi / 3.3 - np.pi / 1.5)) - np.pi / 1.4))
    ys, ncos, mix = [], [], []
    for b in range(nb):
        y, nco = r.pll_combine(sub[b], d["bpf_sq"], zi, 114000, 240000, 0.5, ph, 0.001, st)
        ys.append(y); ncos.append(nco); mix.append(r.fir_mixer(nco, sub[b], d["lpf_3k"], zl)[:-1])
    d["pllc_y"], d["pllc_nco"], d["pllc_state"], d["mixer_y"] = np.stack(ys), np.stack(ncos), st.copy(), np.stack(mix)
    return d


def main():
    m0 = chain_fixture(0, 8, 1, long_blocks=(0, 7))
    np.savez_compressed(os.path.join(HERE, "chain_mode0.npz"), **m0)
    m1 = chain_fixture(1, 3, 2, long_blocks=(0, 2))
    np.savez_compressed(os.path.join(HERE, "chain_mode1.npz"), **m1)
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **func_fixture())
    for f in ("chain_mode0.npz", "chain_mode1.npz", "functions.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
    print(m0["frame_text"])


if __name__ == "__main__":
    main()
