"""Helpers shared by the test modules."""
import hashlib

import numpy as np

F = np.float32
PLL0 = (0.0, 0.0, 1.0, 0.0, 0.0, 1.0)  # src/fm_radio.cpp:165-171
RDS_PHASE = float(F(float(F(np.pi / 3.3 - np.pi / 1.5)) - np.pi / 1.4))  # src/fm_radio.cpp:342,400
LONG_STRIDE = 4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits_equal(a, b):
    """Bit-exact float32 comparison that treats NaN == NaN (payloads ignored) and distinguishes +0/-0."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    if a.dtype != np.float32:
        return bool(np.array_equal(a, b))
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32)))


def assert_bits(a, b, what=""):
    if not bits_equal(a, b):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            raise AssertionError(f"{what}: shape {a.shape} vs {b.shape}")
        bad = np.flatnonzero(~((a == b) | (np.isnan(a.astype(np.float64)) & np.isnan(b.astype(np.float64)))).ravel())
        i = int(bad[0]) if bad.size else -1
        raise AssertionError(f"{what}: {bad.size} of {a.size} differ, first at {i}: {a.ravel()[i]!r} vs {b.ravel()[i]!r}")


def rel_rms(a, b):
    """Relative RMS error ||a-b|| / ||b|| over finite samples (NaN positions must coincide)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN positions differ"
    d = (a - b)[~na]
    den = np.sqrt(np.mean(b[~nb] ** 2))
    return float(np.sqrt(np.mean(d ** 2)) / den) if den > 0 else float(np.sqrt(np.mean(d ** 2)))
