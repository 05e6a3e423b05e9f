"""The reference's Python models (model/fmPll.py, model/fmRRC.py, model/fmSupportLib.py) against the C++-typed operators:
the oracle port on CPU and, with -m gpu, libfmrx through the ctypes binding a model script would use.

The vectors in tests/golden/model.npz were produced by importing the unmodified models (tests/golden/
make_model_golden.py).  The models are a behavioural reference for the C++ program, not a numerical one,
so every tolerance below is stated with its cause; the tight criteria (bit-exact / 1e-5) are checked against the C++
reference elsewhere.  What this file pins down is that the binding lets model/*.py diff outputs directly, and how far
apart the two reference implementations are on the same input."""
import numpy as np
import pytest

from util import rel_rms

F = np.float32
N = 15360


def pilot(n, f, Fs, seed, phase0=0.3, amp=0.08, noise=1e-3):
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    return (amp * np.cos(2 * np.pi * f / Fs * k + phase0) + noise * rng.standard_normal(n)).astype(F)


def fm_iq(n, seed, dev=0.05):
    k = np.arange(n)
    dphi = dev * np.sin(2 * np.pi * k / 97.0) + 0.5 * dev * np.sin(2 * np.pi * k / 31.0 + seed)
    phi = np.cumsum(dphi)
    return np.cos(phi).astype(F), np.sin(phi).astype(F)


@pytest.fixture(scope="module")
def model():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model.npz"))


class OracleOps:
    """the oracle port (CPU restatement of src/*.cpp, pinned against the reference build)"""

    def __init__(self):
        from oracle import Port

        self.p = Port()

    def rrc(self, Fs, n): return self.p.rrc(Fs, n)
    def lpf(self, Fs, Fc, n): return self.p.lpf(Fs, Fc, n)
    def pll(self, x, f, Fs, scale, adj, bw, st): return self.p.pll(x, f, Fs, scale, F(adj), bw, st)
    def demod(self, i, q): return self.p.demod(i, q)
    def fir(self, x, h, zi): return self.p.fir_decim(x, h, zi, 1)


class GpuOps:
    """libfmrx through fmrx/__init__.py (ctypes over the C-ABI)"""

    def __init__(self):
        import fmrx

        self.m = fmrx

    def rrc(self, Fs, n): return self.m.design_rrc(Fs, n)
    def lpf(self, Fs, Fc, n): return self.m.design_lpf(Fs, Fc, n)
    def pll(self, x, f, Fs, scale, adj, bw, st): return self.m.pll(x, f, Fs, scale, float(F(adj)), bw, st).ravel()
    def demod(self, i, q): return self.m.demod(i, q).ravel()
    def fir(self, x, h, zi): return self.m.fir_decim(x, h, zi, 1).ravel()


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def ops(request):
    return OracleOps() if request.param == "oracle" else GpuOps()


def test_rrc_taps_match_fmRRC(ops, model):
    # model/fmRRC.py:11-46 vs src/filter.cpp:63-93: same formula; the C++ stores fp32 and forms t in fp32
    assert rel_rms(ops.rrc(57000.0, 151), model["rrc_57000_151"]) < 2e-7


def test_lpf_taps_match_my_filterImpulseResponse(ops, model):
    # model/fmSupportLib.py:144-154 vs src/filter.cpp:19-38: identical for odd tap counts (for even ones the C++ has the
    # NaN centre tap of Q5, which the model does not); fp32 storage
    for name, (Fc, Fs) in {"lpf_rf": (100e3, 2.4e6), "lpf_mono": (16e3, 240e3), "lpf_3k": (3e3, 240e3)}.items():
        assert rel_rms(ops.lpf(Fs, Fc, 151), model[name]) < 5e-7, name


@pytest.mark.parametrize("name,tol", [("pll_pilot", 2e-3), ("pll_rds", 5e-3)])
def test_pll_tracks_fmPll(ops, model, name, tol):
    """model/fmPll.py:4-56 is float64 throughout; src/helper.cpp:13-57 keeps its state in fp32 and rounds the oscillator
    argument to fp32 (ulp 2e-3 rad at 3e4 rad, Q13), so after two blocks the NCO outputs differ by that quantisation:
    4e-4 (19 kHz, scale 2) and 1.2e-3 (114 kHz) relative RMS — the same loop, locked to the same phase."""
    f, Fs, scale, adj, bw, seed = model[name + "_params"]
    x = pilot(2 * N, f, Fs, int(seed))
    st = np.array([0, 0, 1, 0, 0, 1], F)  # C++ order (src/helper.h:17-19): ..., trigOffset, ncoLast
    nco = np.concatenate([np.asarray(ops.pll(x[b * N:(b + 1) * N], f, Fs, scale, adj, bw, st)).ravel() for b in range(2)])
    err = rel_rms(nco, model[name + "_nco"])
    print(f"{name}: rel-rms vs the float64 model {err:.3g}")
    assert err < tol
    ms = model[name + "_state"]  # model order: integrator, phaseEst, feedbackI, feedbackQ, ncoOut[0], trigOffset
    assert st[4] == ms[5] == 2 * N
    assert abs(st[1] - ms[1]) < 5e-3 and abs(st[2] - ms[2]) < 5e-3 and abs(st[3] - ms[3]) < 5e-3 and abs(st[5] - ms[4]) < 5e-3


def test_discriminator_against_fmDemodArctan(ops, model):
    """model/fmSupportLib.py:12-44 returns the phase step itself (atan2 + unwrap, phase carried across blocks);
    src/rf_module.cpp:13-34 returns (|prev|/|cur|) sin(step) and restarts from (0, 0) at every block (Q3, Q4).  On a
    unit-circle signal with steps <= 0.075 rad they agree to step^2/6: 6e-4 relative RMS, except the first sample of a block,
    which the C++ defines as 0."""
    I, Q = fm_iq(4096, 3)
    dm = np.concatenate([np.asarray(ops.demod(I[b * 2048:(b + 1) * 2048], Q[b * 2048:(b + 1) * 2048])).ravel() for b in range(2)])
    ref = model["demod_atan"]
    mask = np.ones(4096, bool)
    mask[[0, 2048]] = False
    assert dm[0] == 0.0 and dm[2048] == 0.0 and ref[2048] != 0.0
    assert rel_rms(dm[mask], ref[mask]) < 1e-3


def test_convolution_against_my_convoloution(ops, model):
    """model/fmSupportLib.py:157-178 keeps the last 150 inputs as state; src/filter.cpp:126-154 saves them one sample
    late (Q1).  Same arithmetic otherwise (fp32 vs float64): the first block and every output of the second block that no
    longer sees the history agree to 3e-7; the 150 outputs at the block start carry the Q1 glitch."""
    x = np.random.default_rng(5).standard_normal(1024).astype(F)
    h = model["lpf_mono"].astype(F)
    zi = np.zeros(150, F)
    y = np.concatenate([np.asarray(ops.fir(x[b * 512:(b + 1) * 512], h, zi)).ravel() for b in range(2)])
    ref = model["conv_y"]
    assert rel_rms(y[:512], ref[:512]) < 1e-6
    assert rel_rms(y[662:], ref[662:]) < 1e-6
    assert rel_rms(y[512:662], ref[512:662]) > 1e-2  # the one-late state is reproduced, not repaired
