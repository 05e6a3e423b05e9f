"""GPU: every function-level operator of the C-ABI against the golden fixtures (generated from the unmodified
reference) and against the oracle port on fresh random inputs, batched over streams and blocks.

Bars (BASELINE.json north_star): bit-exact for the unpack, for every `exact` FIR / resampler / discriminator and for
the PLL; <= 1e-5 relative RMS for the FMA-rounded RDS stages; bit-exact decoder output.
"""
import numpy as np
import pytest

import fmrx
from oracle import Port, RdsDecoder
from util import F, PLL0, RDS_PHASE, assert_bits, rel_rms

pytestmark = pytest.mark.gpu
TOL = 1e-5  # relative RMS, north_star


@pytest.fixture(scope="module")
def port():
    return Port()


def test_unpack_bit_exact(golden):
    g = golden["functions"]
    assert_bits(fmrx.unpack_iq(g["unpack_in"]), g["unpack_out"], "all 256 byte values")
    rng = np.random.default_rng(1)
    raw = rng.integers(0, 256, 1 << 20, dtype=np.uint8)
    assert_bits(fmrx.unpack_iq(raw), ((raw.astype(np.int32) - 128) / 128.0).astype(F), "1 MiB random")
    assert fmrx.unpack_iq(np.zeros(0, np.uint8)).size == 0


@pytest.mark.parametrize("decim", [1, 5, 10])
def test_fir_decim_golden(golden, decim):
    g = golden["functions"]
    x = g["fir_x"]
    # block by block, state carried by the caller ...
    zi = np.zeros(150, F)
    for b in range(x.shape[0]):
        assert_bits(fmrx.fir_decim(x[b], g["lpf_mono0"], zi, decim), g[f"fir_d{decim}"][b], f"d={decim} block {b}")
    # ... and all blocks in one launch, state carried on the device
    zi2 = np.zeros(150, F)
    assert_bits(fmrx.fir_decim(x, g["lpf_mono0"], zi2, decim), g[f"fir_d{decim}"], f"d={decim} multi-block")
    assert_bits(zi, zi2, "state")
    # FMA rounding stays within tolerance
    zi3 = np.zeros(150, F)
    assert rel_rms(fmrx.fir_decim(x, g["lpf_mono0"], zi3, decim, exact=False), g[f"fir_d{decim}"]) < 1e-6


def test_fir_iq_demod_frontend_golden(golden):
    g = golden["functions"]
    zi, zq = np.zeros(150, F), np.zeros(150, F)
    yi, yq = fmrx.fir_decim_iq(g["fir_x"], g["fir_xq"], g["lpf_rf0"], zi, zq)
    assert_bits(yi, g["fir_iq_i"], "iq.i"); assert_bits(yq, g["fir_iq_q"], "iq.q")
    d = fmrx.demod(yi, yq)
    assert_bits(d, g["demod"], "discriminator")
    assert (d[:, 0] == 0).all()  # Q3: first sample of every block


def test_fir_batched_vs_oracle(port):
    """4 streams x 3 blocks, ragged length (not a multiple of the 1024-output tile), distinct state per stream."""
    rng = np.random.default_rng(7)
    h = fmrx.design_bpf(22e3, 54e3, 240000, 151)
    for decim, n in ((1, 2500), (5, 15360), (10, 10010)):
        x = rng.standard_normal((4, 3, n)).astype(F)
        zi0 = rng.standard_normal((4, 150)).astype(F)
        zi = zi0.copy()
        y = fmrx.fir_decim(x, h, zi, decim)
        for s in range(4):
            z = zi0[s].copy()
            for b in range(3):
                assert_bits(y[s, b], port.fir_decim(x[s, b], h, z, decim), f"d={decim} stream {s} block {b}")
            assert_bits(zi[s], z, "state")


def test_fir_long_state_mode1(port):
    """Mode 1 sizes every state from the 3624-tap filter (src/fm_radio.cpp:189-193): 151-tap FIR with nzi = 3623."""
    rng = np.random.default_rng(8)
    h = fmrx.design_bpf(18.5e3, 19.5e3, 6e6, 151)
    x = rng.standard_normal((2, 15360)).astype(F)
    zi, z = np.zeros(3623, F), np.zeros(3623, F)
    y = fmrx.fir_decim(x, h, zi, 1)
    for b in range(2):
        assert_bits(y[b], port.fir_decim(x[b], h, z, 1), f"block {b}")
    assert_bits(zi, z, "3623-entry state")


def test_frontend_fused_vs_oracle(port):
    """u8 -> unpack -> deinterleave -> FIR/10 (I,Q) -> discriminator in ONE kernel vs the four reference steps."""
    rng = np.random.default_rng(9)
    h = fmrx.design_lpf(2.4e6, 1e5, 151)
    raw = rng.integers(0, 256, (3, 2, 2 * 30720), dtype=np.uint8)  # 3 streams x 2 blocks x 30720 complex samples
    raw[1, 0, :4000] = 128  # a run of exact zeros: I=Q=0 -> zero denominator branch (src/rf_module.cpp:20-23)
    zi, zq = np.zeros((3, 150), F), np.zeros((3, 150), F)
    d, yi, yq = fmrx.frontend(raw, h, zi, zq, want_iq=True)
    for s in range(3):
        a, b = np.zeros(150, F), np.zeros(150, F)
        for k in range(2):
            iq = port.unpack(raw[s, k])
            ri, rq = port.fir_decim_iq(iq[0::2].copy(), iq[1::2].copy(), h, a, b, 10)
            assert_bits(yi[s, k], ri, "I"); assert_bits(yq[s, k], rq, "Q")
            assert_bits(d[s, k], port.demod(ri, rq), f"demod stream {s} block {k}")
        assert_bits(zi[s], a, "zi_i"); assert_bits(zq[s], b, "zi_q")


@pytest.mark.parametrize("n", [30720, 15360 * 10, 30730, 2400])
def test_frontend_state_carry_and_ragged_sizes(port, n):
    """Two calls with the state carried between them (non-zero history at the start of the second call), at sizes that
    take the group-walk kernel + edge kernel (whole runs of whole quads: 30720, the chain's 153600, and 2400 = a single run per block) and a size that takes
    the first form of the kernel (30730: ny = 3073 has no run length and is not a multiple of 4)."""
    rng = np.random.default_rng(n)
    h = fmrx.design_lpf(2.4e6, 1e5, 151)
    S = 2
    zi, zq = np.zeros((S, 150), F), np.zeros((S, 150), F)
    ref = [(np.zeros(150, F), np.zeros(150, F)) for _ in range(S)]
    for call, nb in enumerate((1, 2)):
        raw = rng.integers(0, 256, (S, nb, 2 * n), dtype=np.uint8)
        d, yi, yq = fmrx.frontend(raw, h, zi, zq, want_iq=True)
        for s in range(S):
            a, b = ref[s]
            for k in range(nb):
                iq = port.unpack(raw[s, k])
                ri, rq = port.fir_decim_iq(iq[0::2].copy(), iq[1::2].copy(), h, a, b, 10)
                assert_bits(yi[s, k], ri, f"I call {call} stream {s} block {k}"); assert_bits(yq[s, k], rq, "Q")
                assert_bits(d[s, k], port.demod(ri, rq), f"demod call {call} stream {s} block {k}")
            assert_bits(zi[s], a, "zi_i"); assert_bits(zq[s], b, "zi_q")


def test_resamplers_golden(golden):
    g = golden["functions"]
    cases = [("res_24_125", "lpf_mono1", 125, 24, False, 0, "res_x"), ("res_19_80", "lpf_anti", 80, 19, True, 0, "res_x"),
             ("res_147_800", "lpf_441", 800, 147, False, 0, "res_x441"), ("res_24_5", "lpf_mono1", 5, 24, False, 2949, "res_x")]
    for name, hn, d, u, gain, lim, xn in cases:
        zi = np.zeros(g[hn].size - 1, F)
        y = fmrx.resample(g[xn], g[hn], zi, d, u, gain, lim, exact=True)
        assert_bits(y, g[name], name)
        zi = np.zeros(g[hn].size - 1, F)
        assert rel_rms(fmrx.resample(g[xn], g[hn], zi, d, u, gain, lim, exact=False), g[name]) < 1e-6
    assert np.isnan(g["res_24_125"][:, 12::24]).all()


def test_resample_rejects_short_block():
    with pytest.raises(fmrx.FmrxError, match="longer than the state"):
        fmrx.resample(np.zeros(8000, F), np.zeros(22197, F), np.zeros(22196, F), 800, 147)


def test_pll_golden_bit_exact(golden):
    g = golden["functions"]
    st = np.array(PLL0, F)
    nco = fmrx.pll(g["pll_x"], 19e3, 240e3, 2.0, 0.0, 0.01, st)
    assert_bits(nco, g["pll_nco"], "fmPLL, 3 blocks in one launch")
    assert_bits(st, g["pll_state"], "pll_state_type")


def test_pll_many_lanes_vs_oracle(port):
    """96 streams (three warps' worth, ragged block length) each with its own carrier phase, amplitude and state."""
    rng = np.random.default_rng(11)
    S, B, n = 96, 2, 1999
    k = np.arange(B * n)
    x = np.stack([(0.05 + 0.01 * s) * np.cos(2 * np.pi * 19000 / 240000 * k + 0.1 * s) for s in range(S)]).astype(F)
    x = (x + 0.002 * rng.standard_normal(x.shape)).astype(F).reshape(S, B, n)
    x[5, 0, 100:110] = 0.0  # exact zeros: atan2(+-0, +-0)
    st = np.tile(np.array(PLL0, F), (S, 1))
    nco = fmrx.pll(x, 19e3, 240e3, 2.0, 0.0, 0.01, st)
    for s in range(0, S, 7):
        z = np.array(PLL0, F)
        for b in range(B):
            assert_bits(nco[s, b], port.pll(x[s, b], 19e3, 240e3, 2.0, 0.0, 0.01, z), f"stream {s} block {b}")
        assert_bits(st[s], z, "state")


def test_pll_combine_and_mixer(golden, port):
    """pllCombine (src/helper.cpp:108-173): the filter accumulates DOUBLE products pow(x,2)*h[k] into a float sum, the loop runs
    on that sum -- y, the NCO and the carried state bit-exact against the reference's own output; then the mixer filter
    (src/filter.cpp:373-401, x*x1*h*2 with the half-weight history, Q8), bit-exact too."""
    g = golden["functions"]
    st, zi = np.array(PLL0, F), np.zeros(150, F)
    y, nco = fmrx.pll_combine(g["pllc_x"], g["bpf_sq"], zi, 114000, 240000, 0.5, RDS_PHASE, 0.001, st)
    assert_bits(y, g["pllc_y"], "pllCombine y (double products rounded into a float sum)")
    assert_bits(nco, g["pllc_nco"][:, :-1], "pllCombine NCO")
    assert_bits(st, g["pllc_state"], "state (ncoLast = the reference's untrimmed element)")
    z, zo = np.array(PLL0, F), np.zeros(150, F)
    for b in range(y.shape[0]):
        ry, rn = port.pll_combine(g["pllc_x"][b], g["bpf_sq"], zo, 114000, 240000, 0.5, RDS_PHASE, 0.001, z)
        assert_bits(y[b], ry, f"y vs oracle block {b}"); assert_bits(nco[b], rn[:-1], f"nco vs oracle block {b}")
    assert_bits(zi, zo, "filter state (float squares, one late)")
    zl = np.zeros(150, F)
    m = fmrx.fir_mixer(g["pllc_nco"][:, :-1], g["pllc_x"], g["lpf_3k"], zl)
    assert_bits(m, g["mixer_y"], "mixer filter")


def test_pll_combine_exact_filter_corner_cases(port):
    """The integer rounding of the exact squared-input filter is valid for normal-float partial sums below 4; everything else
    (silence, sums in the subnormal-float range, large inputs, a ragged length that leaves a partial tile, history taps at the
    start of every block) takes the conversion path.  All of it must equal the oracle bit for bit."""
    rng = np.random.default_rng(11)
    h = fmrx.design_bpf(113500.0, 114500.0, 240000.0, 151)
    n = 2500
    cases = {
        "noise": rng.standard_normal((3, n)).astype(F) * F(0.2),
        "large": rng.standard_normal((3, n)).astype(F) * F(40.0),          # sums beyond 4: conversion path
        "tiny": rng.standard_normal((3, n)).astype(F) * F(3e-19),          # squares ~1e-37: partial sums in the subnormal-float range
        "silence_then_noise": np.concatenate([np.zeros((3, 1200), F), rng.standard_normal((3, n - 1200)).astype(F)], 1),
    }
    for name, x in cases.items():
        st, zi = np.array(PLL0, F), np.zeros(150, F)
        y, nco = fmrx.pll_combine(x, h, zi, 114000, 240000, 0.5, RDS_PHASE, 0.001, st)
        z, zo = np.array(PLL0, F), np.zeros(150, F)
        for b in range(3):
            ry, rn = port.pll_combine(x[b], h, zo, 114000, 240000, 0.5, RDS_PHASE, 0.001, z)
            assert_bits(y[b], ry, f"{name}: y block {b}"); assert_bits(nco[b], rn[:-1], f"{name}: nco block {b}")
        assert_bits(zi, zo, name + ": state")


def test_rds_decoder_bit_exact(golden):
    g = golden["chain_mode0"]
    nblk = int(g["nblk"])
    rrc = np.stack([g[f"intent_rds_rrc_{b}"] for b in range(nblk)])
    st = np.zeros((1, fmrx.RDS_STATE_WORDS), np.int32)
    bits, nb, ev, ne = fmrx.rds_decode(rrc, st)
    dec, text = RdsDecoder(), ""
    for b in range(nblk):
        rb, rev = dec.block_decode(rrc[b])
        assert nb[0, b] == rb.size and np.array_equal(bits[0, b, :rb.size], rb), f"bits block {b}"
        assert [tuple(int(v) for v in e) for e in ev[0, b, :ne[0, b]]] == rev, f"events block {b}"
        text += fmrx.rds_format_block(b, int(st[0, 1]), ev[0, b, :ne[0, b]])
    assert text == str(g["frame_text"]) == str(g["binary_frame_text"])
    # same blocks, one call per block with the state carried by the caller, two streams side by side
    st2 = np.zeros((2, fmrx.RDS_STATE_WORDS), np.int32)
    for b in range(nblk):
        b2, n2, e2, m2 = fmrx.rds_decode(np.stack([rrc[b:b + 1], -rrc[b:b + 1]]), st2)
        assert np.array_equal(b2[0, 0], bits[0, b]) and np.array_equal(e2[0, 0, :m2[0, 0]], ev[0, b, :ne[0, b]])
        # an inverted signal decodes to the same bits (differential code) — except the very first one, because the
        # bit prepended in block 0 is the constant front_bit = 0 whatever the polarity (Q12, src/fm_radio.cpp:587-592)
        lo = 1 if b == 0 else 0
        assert np.array_equal(b2[1, 0, lo:n2[1, 0]], bits[0, b, lo:nb[0, b]]), "differential decoding is polarity-blind"


def test_rds_decoder_degenerate_inputs():
    """All-zero and constant input: every comparison ties, stale bits are kept (Q12); must match the oracle."""
    for fill in (0.0, 0.25):
        rrc = np.full((3, 3648), fill, F)
        st = np.zeros((1, fmrx.RDS_STATE_WORDS), np.int32)
        bits, nb, ev, ne = fmrx.rds_decode(rrc, st)
        dec = RdsDecoder()
        for b in range(3):
            rb, rev = dec.block_decode(rrc[b])
            assert np.array_equal(bits[0, b, :nb[0, b]], rb) and [tuple(int(v) for v in e) for e in ev[0, b, :ne[0, b]]] == rev
