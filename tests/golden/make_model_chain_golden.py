"""Generates tests/golden/model_chain.npz: the mono path of the reference's block model (model/fmMonoBlock.py:43-105) and
the stereo carrier recovery (:115-119), run with the model's own functions (scipy.signal.firwin / lfilter as the script
calls them, fmSupportLib.fmDemodArctan, fmPll.fmPll imported from /root/reference/model) on a seeded synthetic
multiplex, block by block with the states carried.  Consumed by tests/test_model_ops.py, which runs the same chain on
the GPU through the fmrx_model_* operators.

    python tests/golden/make_model_chain_golden.py
"""
import os
import sys

import numpy as np
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/model")
sys.path.insert(0, os.path.join(ROOT, "real-time-software-defined-radio_b200"))

from fmPll import fmPll  # noqa: E402
from fmSupportLib import fmDemodArctan  # noqa: E402

from fmrx import synth  # noqa: E402

NBLK, BLOCK = 3, 102400  # the model's block size (fmMonoBlock.py:53): 51200 complex samples


def iq_float(seed=4):
    """the model reads float32 IQ in [-1, 1] (fmMonoBlock.py:39): same multiplex as the u8 generator, not quantised to 8 bits"""
    raw = synth.synth_iq(1, 0, seed=seed)[:NBLK * BLOCK]
    return ((raw.astype(np.float32) - 128.0) / 128.0).astype(np.float32)


def main():
    rf_Fs, rf_Fc, rf_taps, rf_decim = 2.4e6, 100e3, 151, 10
    audio_Fs, audio_Fc, audio_taps, audio_decim = 240e3, 16e3, 151, 5
    iq = iq_float()
    rf_coeff = signal.firwin(rf_taps, rf_Fc / (rf_Fs / 2), window=("hann"))
    audio_coeff = signal.firwin(audio_taps, audio_Fc / (audio_Fs / 2), window=("hann"))
    bp = signal.firwin(rf_taps, [18.5e3 / (audio_Fs / 2), 19.5e3 / (audio_Fs / 2)], window=("hann"), pass_zero="bandpass")
    ext = signal.firwin(rf_taps, [22e3 / (audio_Fs / 2), 54e3 / (audio_Fs / 2)], window=("hann"), pass_zero="bandpass")
    si, sq, sa, sr, se, ss = (np.zeros(rf_taps - 1) for _ in range(6))
    phase = 0.0
    pll_state = [0.0, 0.0, 1.0, 0.0, 1.0, 0.0]
    d = dict(rf_coeff=rf_coeff, audio_coeff=audio_coeff, bp_coeff=bp, nblk=NBLK, block=BLOCK)
    for b in range(NBLK):
        blk = iq[b * BLOCK:(b + 1) * BLOCK]
        i_filt, si = signal.lfilter(rf_coeff, 1.0, blk[0::2], zi=si)
        q_filt, sq = signal.lfilter(rf_coeff, 1.0, blk[1::2], zi=sq)
        i_ds, q_ds = i_filt[::rf_decim], q_filt[::rf_decim]
        fm_demod, phase = fmDemodArctan(i_ds, q_ds, phase)
        audio_filt, sa = signal.lfilter(audio_coeff, 1.0, fm_demod, zi=sa)
        audio_block = audio_filt[::audio_decim]
        bpf, sr = signal.lfilter(bp, 1.0, fm_demod, zi=sr)
        nco, ncoq, pll_state = fmPll(bpf, 19e3, 240e3, pll_state, 2)
        d[f"i_ds_{b}"], d[f"demod_{b}"], d[f"audio_{b}"], d[f"pilot_{b}"], d[f"nco_{b}"] = i_ds, fm_demod, audio_block.copy(), bpf, nco
        # stereo channel extraction, mixing, low-pass, decimation and the combiner exactly as the script writes them
        # (fmMonoBlock.py:150-175) -- including `combined_l_block, combined_r_block = audio_block, audio_block`, which makes
        # all three names one array, so that what the script stores in combined_l AND combined_r is (audio - stereo) / 4
        bpf_ext, se = signal.lfilter(ext, 1.0, fm_demod, zi=se)
        mixed = np.multiply(nco[0:len(bpf_ext):1], bpf_ext) * 2
        stereo_filt, ss = signal.lfilter(audio_coeff, 1.0, mixed, zi=ss)
        stereo_block = stereo_filt[::5]
        combined_l_block, combined_r_block = audio_block, audio_block
        for i in range(len(audio_block)):
            combined_l_block[i] = (audio_block[i] + stereo_block[i]) / 2
            combined_r_block[i] = (audio_block[i] - stereo_block[i]) / 2
        d[f"stereo_{b}"], d[f"combined_{b}"] = stereo_block, combined_l_block.copy()
        assert combined_l_block is combined_r_block
    d["phase"], d["pll_state"] = phase, np.array(pll_state)
    # the RDS resampler's arithmetic (fmRDSblock.py:188-199): zero-stuff by 19, anti-image lfilter, [::80], x19
    rng = np.random.default_rng(8)
    x = rng.standard_normal(2 * 960)
    anti = signal.firwin(151, (57000 / 2) / ((240000 * 19) / 2), window=("hann"))
    st, outs = np.zeros(150), []
    for b in range(2):
        upx = np.zeros(960 * 19)
        upx[::19] = x[b * 960:(b + 1) * 960]
        y, st = signal.lfilter(anti, 1.0, upx, zi=st)
        outs.append(y[::80] * 19)
    d["anti_coeff"], d["resampled"] = anti, np.concatenate(outs)
    np.savez_compressed(os.path.join(HERE, "model_chain.npz"), **d)
    print({k: np.shape(v) for k, v in d.items()})


if __name__ == "__main__":
    main()
