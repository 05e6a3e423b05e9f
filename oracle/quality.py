"""TEST INFRASTRUCTURE ONLY.  Independent restatement (numpy / scipy, float64) of the `quality` profile's building blocks
(include/fmrx.h FMRX_QUALITY_*).  None of this exists in the reference -- its report proposes a de-emphasis
filter, unity-gain band-pass filters and an automatic RDS phase adjust and the program never got them -- so the oracle here is
the textbook definition, not a reference function: parity status "no reference; pinned by definition" (scipy.signal.bilinear /
lfilter for the de-emphasis, a direct DFT sum for filter responses, an exhaustive phase sweep through the reference's own
functions for the phase adjust)."""
import numpy as np
from scipy import signal


def deemphasis_ba(tau_us, fs):
    """1 / (1 + s tau) through the bilinear transform"""
    return signal.bilinear([1.0], [tau_us * 1e-6, 1.0], fs)


def deemphasis(x, tau_us, fs, zi=None):
    """lfilter along the last axis; returns (y, zf) with scipy's state"""
    b, a = deemphasis_ba(tau_us, fs)
    if zi is None:
        zi = np.zeros(np.shape(x)[:-1] + (1,))
    return signal.lfilter(b, a, np.asarray(x, np.float64), zi=zi)


def response(h, fs, f):
    k = np.arange(len(h))
    z = np.sum(np.asarray(h, np.float64) * np.exp(-2j * np.pi * f * k / fs))
    return abs(z), float(np.angle(z))
