"""Generates tests/golden/model.npz by IMPORTING the reference's Python models (model/fmPll.py, model/fmRRC.py,
model/fmSupportLib.py under /root/reference) in the authoring container and running them on seeded inputs.

    python tests/golden/make_model_golden.py

The models are a BEHAVIOURAL reference for the C++ program, not a numerical one (float64 throughout, a
true atan2 discriminator with carried phase, correct FIR state, different filter design for even tap counts), so the
tests that consume these vectors (tests/test_model_parity.py) state, operator by operator, how close the C++-typed
operators of libfmrx / the oracle are to them and why.  Inputs are regenerated from the seeds at test time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODEL = "/root/reference/model"
sys.path.insert(0, MODEL)

import fmPll  # noqa: E402
import fmRRC  # noqa: E402
import fmSupportLib  # noqa: E402


def pilot(n, f, Fs, seed, phase0=0.3, amp=0.08, noise=1e-3):
    rng = np.random.default_rng(seed)
    k = np.arange(n)
    return (amp * np.cos(2 * np.pi * f / Fs * k + phase0) + noise * rng.standard_normal(n)).astype(np.float32)


def fm_iq(n, seed, dev=0.05):
    """unit-circle IQ whose phase advances by dev*sin(.) rad per sample (small deviation: the two discriminators agree)"""
    k = np.arange(n)
    dphi = dev * np.sin(2 * np.pi * k / 97.0) + 0.5 * dev * np.sin(2 * np.pi * k / 31.0 + seed)
    phi = np.cumsum(dphi)
    return np.cos(phi).astype(np.float32), np.sin(phi).astype(np.float32)


def main():
    d = {}
    # ---- fmRRC.py:11-46 and fmSupportLib.py:144-154
    d["rrc_57000_151"] = fmRRC.impulseResponseRootRaisedCosine(57000, 151)
    for name, (Fc, Fs) in {"lpf_rf": (100e3, 2.4e6), "lpf_mono": (16e3, 240e3), "lpf_3k": (3e3, 240e3)}.items():
        d[name] = fmSupportLib.my_filterImpulseResponse(Fc, Fs, 151)
    # ---- fmPll.py:4-56, two consecutive blocks with the state carried, stereo and RDS settings of fmMonoBlock.py / fmRDSblock.py
    n = 15360
    for name, (f, scale, adj, bw, seed) in {"pll_pilot": (19e3, 2.0, 0.0, 0.01, 11), "pll_rds": (114e3, 0.5, np.pi / 3.3 - np.pi / 1.5, 0.001, 12)}.items():
        x = pilot(2 * n, f, 240e3, seed)
        st = [0.0, 0.0, 1.0, 0.0, 1.0, 0.0]  # model order: integrator, phaseEst, feedbackI, feedbackQ, ncoOut[0], trigOffset
        outs = []
        for b in range(2):
            nco, ncoq, st = fmPll.fmPll(x[b * n:(b + 1) * n].astype(np.float64), f, 240e3, st, ncoScale=scale, phaseAdjust=adj, normBandwidth=bw)
            outs.append(nco[:-1])
        d[name + "_nco"] = np.concatenate(outs)
        d[name + "_state"] = np.array(st)
        d[name + "_params"] = np.array([f, 240e3, scale, adj, bw, seed], np.float64)
    # ---- fmSupportLib.py:12-44, two blocks with the phase carried
    I, Q = fm_iq(4096, 3)
    prev, outs = 0.0, []
    for b in range(2):
        dm, prev = fmSupportLib.fmDemodArctan(I[b * 2048:(b + 1) * 2048], Q[b * 2048:(b + 1) * 2048], prev)
        outs.append(dm)
    d["demod_atan"] = np.concatenate(outs)
    # ---- fmSupportLib.py:157-178 my_convoloution, two blocks with its (correct, not one-late) state
    rng = np.random.default_rng(5)
    x = rng.standard_normal(1024).astype(np.float32)
    h = d["lpf_mono"].astype(np.float32)
    zi, outs = np.zeros(150), []
    for b in range(2):
        y, zi = fmSupportLib.my_convoloution(x[b * 512:(b + 1) * 512].astype(np.float64), h.astype(np.float64), 151, zi)
        outs.append(y)
    d["conv_y"] = np.concatenate(outs)
    np.savez_compressed(os.path.join(HERE, "model.npz"), **d)
    print({k: v.shape for k, v in d.items()})


if __name__ == "__main__":
    main()
