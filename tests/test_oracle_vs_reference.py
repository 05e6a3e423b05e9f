"""CPU: the oracle port against the UNMODIFIED reference, live — the reference objects behind oracle/ref_shim.cpp
(oracle/_ref/libfmref.so) and the reference executable (oracle/_ref/fm_radio), both built by oracle/Makefile straight
from /root/reference/src.  This is what pins the oracle; it runs wherever oracle/_ref has been built (the authoring
container, and the GPU box, to which the built files travel) and is skipped elsewhere — tests/test_oracle_golden.py
covers the same ground there from committed fixtures.  Everything is bit-exact."""
import numpy as np
import pytest

from fmrx import synth
from oracle import Chain, Port
from oracle.ref import Ref, RefChain, ref_available, run_ref_binary
from util import F, PLL0, RDS_PHASE, assert_bits

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.fixture(scope="module")
def pair():
    return Port(), Ref()


def test_filter_designs(pair):
    port, ref = pair
    for Fs, Fc, n in ((2.4e6, 1e5, 151), (240000, 16000, 151), (6e6, 16000, 3624), (240000, 3000, 151), (4560000, 28500, 2869)):
        assert_bits(port.lpf(Fs, Fc, n), ref.lpf(Fs, Fc, n), f"lpf {Fs} {Fc} {n}")
    for Fb, Fe in ((18.5e3, 19.5e3), (22e3, 54e3), (54000, 60000), (113500, 114500)):
        assert_bits(port.bpf(Fb, Fe, 240000, 151), ref.bpf(Fb, Fe, 240000, 151), f"bpf {Fb}-{Fe}")
    assert_bits(port.rrc(57000, 151), ref.rrc(57000, 151), "rrc")


def test_unpack_and_front_end(pair):
    port, ref = pair
    rng = np.random.default_rng(3)
    raw = rng.integers(0, 256, 2 * 4000, dtype=np.uint8)
    assert_bits(port.unpack(raw), ref.unpack(raw), "unpack")
    h = ref.lpf(2.4e6, 1e5, 151)
    xi, xq = port.unpack(raw)[0::2].copy(), port.unpack(raw)[1::2].copy()
    zp = [np.zeros(150, F), np.zeros(150, F)]
    zr = [np.zeros(150, F), np.zeros(150, F)]
    for _ in range(3):  # carried state, one-late history (Q1)
        yp = port.fir_decim_iq(xi, xq, h, zp[0], zp[1], 10)
        yr = ref.fir_decim_iq(xi, xq, h, zr[0], zr[1], 10)
        assert_bits(yp[0], yr[0], "I"); assert_bits(yp[1], yr[1], "Q")
        assert_bits(zp[0], zr[0], "zi I"); assert_bits(zp[1], zr[1], "zi Q")
        assert_bits(port.demod(yp[0], yp[1]), ref.demod(yr[0], yr[1]), "demod")


@pytest.mark.parametrize("decim", [1, 5])
def test_fir_decim(pair, decim):
    port, ref = pair
    rng = np.random.default_rng(4)
    h = ref.bpf(22e3, 54e3, 240000, 151)
    zp, zr = np.zeros(150, F), np.zeros(150, F)
    for _ in range(3):
        x = rng.standard_normal(3000).astype(F)
        assert_bits(port.fir_decim(x, h, zp, decim), ref.fir_decim(x, h, zr, decim), "y")
        assert_bits(zp, zr, "zi")


@pytest.mark.parametrize("up,decim,taps,rds", [(24, 125, 3624, False), (19, 80, 2869, True), (147, 800, 22197, False)])
def test_resamplers(pair, up, decim, taps, rds):
    port, ref = pair
    rng = np.random.default_rng(5)
    h = ref.lpf(240000.0 * up, 16000, taps)
    zp, zr = np.zeros(taps - 1, F), np.zeros(taps - 1, F)
    for _ in range(2):
        x = rng.standard_normal(taps + 400).astype(F)
        yp = port.resample(x, h, zp, decim, up, gain_up=rds)
        yr = ref.resample_rds(x, h, zr, decim, up) if rds else ref.resample_ptr(x, h, zr, decim, up)
        assert_bits(yp, yr, "y")
        assert_bits(zp, zr, "zi")


def test_plls(pair):
    port, ref = pair
    n = 15360
    t = np.arange(2 * n) / 240000.0
    for freq, scale, adj, bw in ((19e3, 2.0, 0.0, 0.01), (114000.0, 0.5, RDS_PHASE, 0.001)):
        x = (0.05 * np.cos(2 * np.pi * freq * t + 0.4)).astype(F)
        sp, sr = np.array(PLL0, F), np.array(PLL0, F)
        for b in range(2):
            assert_bits(port.pll(x[b * n:(b + 1) * n], freq, 240000.0, scale, adj, bw, sp),
                        ref.pll(x[b * n:(b + 1) * n], freq, 240000.0, scale, adj, bw, sr), f"nco {freq} block {b}")
            assert_bits(sp, sr, "pll state")


@pytest.mark.parametrize("mode", [0, 1])
def test_chain_vs_reference_functions_and_binary(mode):
    """Five blocks through (a) the oracle port's chain, (b) the reference FUNCTIONS composed as fm_radio.cpp's thread
    bodies compose them, (c) the reference EXECUTABLE: int16 audio of (a) == (b) in both profiles, == (c) in the
    `binary` profile; RDS sync lines of (a) == the executable's stderr."""
    nblk = 5
    raw = synth.synth_iq(nblk, mode, seed=7)
    out_bin, err_text, rc = run_ref_binary(raw, mode)
    assert rc == 0
    for profile in (0, 1):
        audio, _, bits, events, _ = Chain(mode, profile).run(raw)
        rc_chain = RefChain(mode, profile)
        ref_audio = np.concatenate([rc_chain.block(raw[b * 307200:(b + 1) * 307200]) for b in range(nblk)])
        assert_bits(audio, ref_audio, f"mode {mode} profile {profile}: port vs reference functions")
        if profile == 0:
            n = audio.size  # whole input blocks only: the executable may run one more block on the short read at EOF (Q9)
            assert_bits(audio, out_bin[:n], f"mode {mode}: port vs the reference executable's stdout")
    if mode == 0:
        for _, kind, letter, pos in events:
            line = "~~~~~Re-Sync~~~~~" if kind == 2 else ("False positive " if kind == 1 else "") + f"Syndrome {'ABCD'[letter]} at position {pos}"
            assert line in err_text, line
