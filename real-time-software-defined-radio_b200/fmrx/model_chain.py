"""The block loop of the reference's Python model (model/fmMonoBlock.py:43-175) as one object over the GPU's
model-compatible operators (fmrx_model_firwin / lfilter / demod / pll, csrc/fmrx_model.cu), so that the
script's arrays can be diffed directly: float64 throughout, scipy's Hann firwin taps, lfilter states, arctan
discriminator with the carried phase, fmPll with the model's state order, the x2 on the stereo mixer.

Batched: `S` independent stations step together, one block of float IQ per station per call.  States live in this object
(numpy, float64) and are updated in place by the operators, exactly as the script carries them from block to block.
"""
import numpy as np

from . import model_demod, model_firwin, model_lfilter, model_pll


class ModelMonoStereo:
    # fmMonoBlock.py:43-52
    RF_FS, RF_FC, RF_TAPS, RF_DECIM = 2.4e6, 100e3, 151, 10
    AUDIO_FS, AUDIO_FC, AUDIO_TAPS, AUDIO_DECIM = 240e3, 16e3, 151, 5

    def __init__(self, n_streams=1):
        S, nyq = n_streams, self.AUDIO_FS / 2
        self.S = S
        self.rf_coeff = model_firwin(self.RF_TAPS, self.RF_FC / (self.RF_FS / 2))                                 # :55
        self.audio_coeff = model_firwin(self.AUDIO_TAPS, self.AUDIO_FC / nyq)                                     # :58
        self.recovery_coeff = model_firwin(self.RF_TAPS, [18.5e3 / nyq, 19.5e3 / nyq], pass_zero=False)            # :115
        self.extraction_coeff = model_firwin(self.RF_TAPS, [22e3 / nyq, 54e3 / nyq], pass_zero=False)              # :151
        self.stereo_coeff = model_firwin(self.RF_TAPS, self.AUDIO_FC / nyq)                                       # :160
        z = lambda n: np.zeros((S, n))  # noqa: E731
        self.state_i, self.state_q = z(self.RF_TAPS - 1), z(self.RF_TAPS - 1)                                     # :47-48
        self.state_audio, self.state_recovery = z(self.AUDIO_TAPS - 1), z(self.RF_TAPS - 1)
        self.state_extraction, self.state_stereo = z(self.RF_TAPS - 1), z(self.AUDIO_TAPS - 1)
        self.state_phase = np.zeros(S)                                                                            # :49
        self.recovery_state = np.tile(np.array([0.0, 0.0, 1.0, 0.0, 1.0, 0.0]), (S, 1))                           # :75

    def block(self, iq):
        """iq: float [S][2n] (or [2n] for one station) interleaved I,Q in [-1, 1].  Returns a dict of the script's per-block
        arrays: i_ds, q_ds, fm_demod, audio (mono), pilot, nco, stereo, left / right ((audio +- stereo) / 2, what the
        combiner means) and combined (what the script's combiner stores in BOTH combined_l and combined_r: its three
        names alias one array, :166-170, so the stored value is (audio - stereo) / 4)."""
        iq = np.asarray(iq, np.float64).reshape(self.S, -1)
        i_ds = model_lfilter(iq[:, 0::2], self.rf_coeff, self.state_i, decim=self.RF_DECIM)                       # :86-95
        q_ds = model_lfilter(iq[:, 1::2], self.rf_coeff, self.state_q, decim=self.RF_DECIM)
        fm_demod = model_demod(i_ds, q_ds, self.state_phase)                                                      # :98
        audio = model_lfilter(fm_demod, self.audio_coeff, self.state_audio, decim=self.AUDIO_DECIM)               # :101-105
        pilot = model_lfilter(fm_demod, self.recovery_coeff, self.state_recovery)                                 # :117
        nco, _ = model_pll(pilot, 19e3, 240e3, self.recovery_state, 2.0)                                          # :119
        extracted = model_lfilter(fm_demod, self.extraction_coeff, self.state_extraction)                         # :152
        mixed = nco.reshape(self.S, -1)[:, :extracted.reshape(self.S, -1).shape[1]] * extracted.reshape(self.S, -1) * 2  # :156-157
        stereo = model_lfilter(mixed, self.stereo_coeff, self.state_stereo, decim=5)                              # :161-163
        a, s = np.asarray(audio).reshape(self.S, -1), np.asarray(stereo).reshape(self.S, -1)
        return {"i_ds": i_ds, "q_ds": q_ds, "fm_demod": fm_demod, "audio": audio, "pilot": pilot, "nco": nco, "stereo": stereo,
                "left": (a + s) / 2, "right": (a - s) / 2, "combined": ((a + s) / 2 - s) / 2}
