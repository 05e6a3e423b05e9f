"""SURVEY 8(f) row 4, the `quality` profile (include/fmrx.h FMRX_QUALITY_*): de-emphasis, unity-gain band-pass filters with the
x2 stereo mixer and the delayed mono branch, computed RDS phase adjust.  None of it is in the reference, so the checks are: the
building blocks against their textbook definition (oracle/quality.py: scipy.signal.bilinear / lfilter, DFT sums), the GPU chain
against the oracle chain with the same flags, and that each flag does what it is for -- measured on synthetic input: stereo
separation, audio level of a pre-emphasised tone, RDS symbol amplitude over a phase sweep.  The default (quality = 0) stays the
reference's receiver: every other test file runs with it."""
import numpy as np
import pytest

import fmrx
from fmrx import synth
from oracle import Chain
from oracle.quality import deemphasis, deemphasis_ba, response
from util import F, assert_bits, rel_rms

Q75, Q50, QU, QP = fmrx.QUALITY_DEEMPH_75, fmrx.QUALITY_DEEMPH_50, fmrx.QUALITY_UNITY_BPF, fmrx.QUALITY_AUTO_RDS_PHASE


def test_design_helpers_against_their_definitions():
    """host code only (no GPU): coefficients, unity gain, phase formula"""
    for tau, fs in ((75.0, 48000.0), (50.0, 48000.0), (75.0, 44100.0)):
        b, a1 = fmrx.deemphasis_coeffs(tau, fs)
        bb, aa = deemphasis_ba(tau, fs)
        assert abs(b - bb[0]) < 1e-15 and abs(b - bb[1]) < 1e-15 and abs(a1 - aa[1]) < 1e-15
        w, h = __import__("scipy.signal", fromlist=["freqz"]).freqz(bb, aa, worN=[2 * np.pi * 2122.0 / fs])  # 1 / (2 pi 75 us) = 2122 Hz: -3 dB
        if tau == 75.0:
            assert abs(20 * np.log10(abs(h[0])) + 3.01) < 0.05
    for fb, fe in ((18.5e3, 19.5e3), (22e3, 54e3), (54e3, 60e3), (113.5e3, 114.5e3)):
        h0, h1 = fmrx.design_bpf(fb, fe, 240e3, 151), fmrx.design_bpf_unity(fb, fe, 240e3, 151)
        g0, g1 = response(h0, 240e3, (fb + fe) / 2)[0], response(h1, 240e3, (fb + fe) / 2)[0]
        assert abs(g1 - 1.0) < 1e-6 and abs(fmrx.fir_response(h1, 240e3, (fb + fe) / 2)[0] - g1) < 1e-12
        assert np.allclose(h1, h0 / g0, rtol=1e-6, atol=0)
    assert 0.30 < response(fmrx.design_bpf(18.5e3, 19.5e3, 240e3, 151), 240e3, 19e3)[0] < 0.32, "the reference's pilot band-pass has a gain of 0.308"
    hq = fmrx.design_bpf(113.5e3, 114.5e3, 240e3, 151)
    assert abs(fmrx.rds_auto_phase(hq) + 0.5 * response(hq, 240e3, 114e3)[1]) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("tau,n", [(75.0, 3072), (50.0, 2949), (75.0, 2822), (75.0, 37)])
def test_deemphasis_operator_vs_scipy(tau, n):
    """the blocked-scan kernel (fp32) against scipy.signal.lfilter (float64) with the state carried over two calls of two blocks;
    a NaN input sample counts as 0 (the quantiser's rule, src/fm_radio.cpp:290-293)"""
    rng = np.random.default_rng(int(tau) + n)
    S, B = 5, 2
    fs = 48000.0
    state, zi = np.zeros((S, 4), F), None
    for call in range(2):
        x = (0.3 * rng.standard_normal((S, B, 2 * n))).astype(F)
        if call == 1:
            x[2, 0, 10] = np.nan
        y, q = fmrx.deemphasis(x, tau, fs, state, mult=1, want_int16=True)
        xs = np.nan_to_num(x.astype(np.float64), nan=0.0).reshape(S, B * n, 2).transpose(0, 2, 1)   # [S][ch][time]
        ref, zi = deemphasis(xs, tau, fs, zi)
        ref = ref.transpose(0, 2, 1).reshape(S, B, 2 * n)
        err = rel_rms(y, ref)
        print(f"tau {tau} n {n} call {call}: rel-rms vs lfilter {err:.3g}")
        assert err < 1e-6
        qi = np.trunc(ref.astype(F) * F(16384.0)).astype(np.int64)
        assert np.abs(q.astype(np.int64) - qi).max() <= 1


@pytest.mark.gpu
@pytest.mark.parametrize("mode,quality", [(0, Q75 | QU | QP), (0, QU), (0, Q50), (1, Q75 | QU), (2, Q75 | QU | QP)])
def test_quality_chain_vs_oracle_chain(mode, quality):
    """GPU chain with quality flags against the oracle chain with the same flags: everything ahead of the de-emphasis bit-exact
    (same taps, same roundings), the de-emphasised float audio <= 1e-5 relative RMS, int16 +-1 LSB, RDS bits and events equal."""
    S, B = 3, 4
    raw = np.stack([synth.synth_station(s * 7, B, mode) for s in range(S)])
    with fmrx.Batch(S, mode=mode, profile=1, max_blocks=2, quality=quality) as rx:
        res = [rx.process(raw[:, :2 * 307200], want_float=True), rx.process(raw[:, 2 * 307200:], want_float=True)]
        mono = None
    audio_f = np.concatenate([r["audio_f"] for r in res], 1)
    audio = np.concatenate([r["audio"] for r in res], 1)
    for s in range(S):
        ch = Chain(mode, 1, quality=quality)
        ref_i16, cap, bits, events, _ = ch.run(raw[s], taps=("audio_f",))
        ref_f = np.stack(cap["audio_f"])
        if quality & (Q75 | Q50):
            err = rel_rms(np.nan_to_num(audio_f[s]), np.nan_to_num(ref_f))
            print(f"mode {mode} quality {quality} station {s}: float audio rel-rms {err:.3g}")
            assert err < 1e-5
            assert np.abs(audio[s].ravel().astype(int) - ref_i16.astype(int)).max() <= 1
        else:
            assert_bits(audio_f[s], ref_f, f"station {s} float audio (no de-emphasis: bit-exact)")
            assert_bits(audio[s].ravel(), ref_i16, f"station {s} int16")
        if mode != 1:
            got = np.concatenate([r["rds_bits"][s, b, :r["rds_n_bits"][s, b]] for r in res for b in range(2)])
            assert np.array_equal(got, np.concatenate(bits)), f"station {s} RDS bits"


@pytest.mark.gpu
def test_quality_profile_does_what_it_is_for():
    """Measured on synthetic input: (1) a left-only tone -- the reference's receiver leaks it into R at about the level of L
    (L-R arrives at a third of the gain of L+R and 0.31 ms late); with UNITY_BPF the separation is better than 25 dB; (2) RDS
    symbol energy over a sweep of the NCO phase adjust: the computed adjust is within 0.5 % of the best of the sweep and above
    the hand-tuned constant; (3) de-emphasis: a 10 kHz tone comes out 13.7 dB (75 us) below a 400 Hz tone of the same deviation."""
    B = 6
    t = np.arange(B * 153600) / 2.4e6

    def synth_lr(fl, fr, al=0.5, ar=0.0):
        left, right = al * np.sin(2 * np.pi * fl * t), ar * np.sin(2 * np.pi * fr * t)
        th = 2 * np.pi * 19000.0 * t
        m = 0.45 * (left + right) + 0.45 * (left - right) * np.cos(2 * th) + 0.08 * np.cos(th)
        phi = 2 * np.pi * 75e3 * np.cumsum(m) / 2.4e6
        out = np.empty(2 * t.size, np.uint8)
        out[0::2] = np.clip(np.rint(127.0 * np.cos(phi) + 128.0), 0, 255)
        out[1::2] = np.clip(np.rint(127.0 * np.sin(phi) + 128.0), 0, 255)
        return out

    def rms_at(x, f):  # amplitude of the component at f in the last blocks of a 48 kHz signal
        x = x[-3 * 3072:]
        n = np.arange(x.size)
        return 2 * abs(np.sum(x * np.exp(-2j * np.pi * f * n / 48000.0))) / x.size

    raw = synth_lr(1000.0, 3000.0)
    sep = {}
    for name, q in (("reference", 0), ("quality", QU)):
        with fmrx.Batch(1, mode=0, profile=1, max_blocks=B, quality=q) as rx:
            a = rx.process(raw, want_float=True)["audio_f"][0].reshape(-1, 2)
        sep[name] = 20 * np.log10(rms_at(a[:, 0], 1000.0) / max(rms_at(a[:, 1], 1000.0), 1e-12))
    print(f"stereo separation of a left-only 1 kHz tone: reference receiver {sep['reference']:.1f} dB, UNITY_BPF {sep['quality']:.1f} dB")
    assert sep["reference"] < 6.0 and sep["quality"] > 25.0
    # (2) phase sweep
    raw = synth.synth_iq(B, 0, seed=3)
    def symbol_rms(phase=None, quality=0):
        with fmrx.Batch(1, mode=0, profile=1, max_blocks=B, paths=fmrx.PATH_RDS | fmrx.PATH_RDS_STAGES, quality=quality) as rx:
            if phase is not None:
                rx.rds_phase = phase
            rx.process(raw)
            used = rx.rds_phase
            rrc = rx.tap("rds_rrc")[0, 2:]
        return float(np.sqrt(np.mean(rrc.astype(np.float64) ** 2))), used
    sweep = [symbol_rms(p)[0] for p in np.linspace(-np.pi / 2, np.pi / 2, 25)]
    hand, hand_phase = symbol_rms()
    auto, auto_phase = symbol_rms(quality=QP)
    print(f"RDS RRC output rms: hand-tuned constant ({hand_phase:.4f} rad) {hand:.4f}, computed adjust ({auto_phase:.4f} rad) {auto:.4f}, best of a 25-point sweep {max(sweep):.4f}")
    assert auto >= 0.995 * max(sweep) and auto > hand
    # (3) de-emphasis response on demodulated tones
    lv = {}
    for f in (400.0, 10000.0):
        raw = synth_lr(f, f, al=0.5, ar=0.5)
        with fmrx.Batch(1, mode=0, profile=1, max_blocks=B, quality=Q75) as rx:
            a = rx.process(raw, want_float=True)["audio_f"][0].reshape(-1, 2)
        with fmrx.Batch(1, mode=0, profile=1, max_blocks=B) as rx:
            a0 = rx.process(raw, want_float=True)["audio_f"][0].reshape(-1, 2)
        lv[f] = 20 * np.log10(rms_at(a[:, 0], f) / rms_at(a0[:, 0], f))
    want = {f: -10 * np.log10(1 + (2 * np.pi * f * 75e-6) ** 2) for f in lv}
    print(f"de-emphasis 75 us: {lv[400.0]:.2f} dB at 400 Hz ({want[400.0]:.2f} analogue), {lv[10000.0]:.2f} dB at 10 kHz ({want[10000.0]:.2f} analogue; the bilinear transform warps the top of the band)")
    assert abs(lv[400.0] - want[400.0]) < 0.1 and -16.5 < lv[10000.0] < -13.0
