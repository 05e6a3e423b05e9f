"""fmrx — ctypes binding of libfmrx.so (include/fmrx.h), the B200-native FM receive chain.

Two layers:
  * `fmrx.lib()` / the thin wrappers below: one Python function per C-ABI entry, numpy arrays in and out;
  * `fmrx.refapi`: the same operators under the reference's own function names and argument order
    (impulseResponseLPF, convolveWithDecim, fmPLL, ... — /root/reference/src/filter.h, helper.h, rf_module.h), so a
    test written against the reference reads the same against this library.

There is no CPU fallback: importing works anywhere (so the not-gpu tests can check the ABI), but every compute call
raises FmrxError when the extension is missing or no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_DIR, "libfmrx.so")
CLI_PATH = os.path.join(PKG_DIR, "fm_radio")

BLOCK_BYTES = 307200
IF_PER_BLOCK = 15360
RDS_PER_BLOCK = 3648
MAX_EVENTS = 160
MAX_BITS = 80
RDS_STATE_WORDS = 160
PROFILE_BINARY, PROFILE_INTENT = 0, 1
PATH_AUDIO, PATH_RDS, PATH_RDS_STAGES = 1, 2, 4
NUMERICS_REFERENCE, NUMERICS_FMA, NUMERICS_STRICT = 0, 1, 2
QUALITY_DEEMPH_75, QUALITY_DEEMPH_50, QUALITY_UNITY_BPF, QUALITY_AUTO_RDS_PHASE = 1, 2, 4, 8
TAPS = dict(demod=0, mono=1, pilot=2, nco=3, stereo_bpf=4, stereo=5, rds_bpf=6, rds_sq=7, rds_nco=8, rds_lpf=9, rds_res=10, rds_rrc=11)

STAGES = ("frontend", "mono", "pilot_bpf", "stereo_bpf", "rds_bpf", "rds_sq_bpf", "pll", "stereo_lpf", "combine", "rds_mix_lpf",
          "rds_resample", "rds_rrc", "rds_decode", "rds_symbols", "bpf_fused")

F = np.float32
fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
dp = C.POINTER(C.c_double)


class RdsGroup(C.Structure):
    _fields_ = [("blk", C.c_uint16 * 4), ("type", C.c_uint8), ("version_b", C.c_uint8), ("corrected", C.c_uint8), ("reserved", C.c_uint8), ("bit_index", C.c_uint32)]


class RdsStation(C.Structure):
    _fields_ = [("synced", C.c_int32), ("pi", C.c_int32), ("pty", C.c_int32), ("tp", C.c_int32), ("ps", C.c_char * 9), ("rt", C.c_char * 65),
                ("ps_complete", C.c_uint8), ("rt_ab_flag", C.c_uint8), ("groups", C.c_uint32), ("blocks_ok", C.c_uint32), ("blocks_corrected", C.c_uint32),
                ("blocks_bad", C.c_uint32), ("sync_losses", C.c_uint32), ("bits_fed", C.c_uint64)]


class FmrxError(RuntimeError):
    pass


class RdsEvent(C.Structure):
    _fields_ = [("block", C.c_int32), ("kind", C.c_int32), ("letter", C.c_int32), ("position", C.c_uint32)]


EVENT_DTYPE = np.dtype([("block", np.int32), ("kind", np.int32), ("letter", np.int32), ("position", np.uint32)])
evp = C.POINTER(RdsEvent)


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("mode", "profile", "n_streams", "max_blocks", "device", "paths", "numerics", "quality")]


class Outputs(C.Structure):
    _fields_ = [("audio", i16p), ("audio_f", fp), ("rds_bits", u8p), ("rds_n_bits", i32p), ("rds_events", evp), ("rds_n_events", i32p)]


# every exported symbol with its signature — tests/test_abi.py checks this table against include/fmrx.h
SIGNATURES = {
    "fmrx_last_error": (C.c_char_p, []),
    "fmrx_version": (C.c_int, []),
    "fmrx_device_count": (C.c_int, []),
    "fmrx_design_lpf": (C.c_int, [C.c_float, C.c_float, C.c_ushort, fp]),
    "fmrx_design_bpf": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_int, fp]),
    "fmrx_design_rrc": (C.c_int, [C.c_float, C.c_int, fp]),
    "fmrx_fir_response": (C.c_int, [fp, C.c_int, C.c_float, C.c_float, dp, dp]),
    "fmrx_design_bpf_unity": (C.c_int, [C.c_float, C.c_float, C.c_float, C.c_int, fp]),
    "fmrx_rds_auto_phase": (C.c_int, [fp, C.c_int, C.c_float, C.c_float, fp]),
    "fmrx_deemphasis_coeffs": (C.c_int, [C.c_float, C.c_float, dp, dp]),
    "fmrx_deemphasis": (C.c_int, [fp, i16p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, fp]),
    "fmrx_batch_rds_phase": (C.c_float, [C.c_void_p]),
    "fmrx_batch_set_rds_phase": (C.c_int, [C.c_void_p, C.c_float]),
    "fmrx_unpack_iq": (C.c_int, [u8p, C.c_size_t, fp]),
    "fmrx_fir_decim": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int]),
    "fmrx_fir_decim_iq": (C.c_int, [fp, fp, fp, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, fp, C.c_int, C.c_int]),
    "fmrx_resample": (C.c_int, [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fmrx_fir_mixer": (C.c_int, [fp, fp, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp]),
    "fmrx_demod": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, fp]),
    "fmrx_pll": (C.c_int, [fp, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]),
    "fmrx_pll_combine": (C.c_int, [fp, fp, fp, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]),
    "fmrx_frontend": (C.c_int, [fp, fp, fp, u8p, C.c_int, C.c_int, C.c_int, fp, C.c_int, fp, fp, C.c_int]),
    "fmrx_rds_decode": (C.c_int, [fp, C.c_int, C.c_int, C.c_int, u8p, i32p, evp, i32p, i32p]),
    "fmrx_rds_state_offset": (C.c_int, [i32p]),
    "fmrx_rds_format_block": (C.c_int, [C.c_int, C.c_int, evp, C.c_int, C.c_char_p, C.c_int]),
    "fmrx_batch_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "fmrx_batch_destroy": (None, [C.c_void_p]),
    "fmrx_batch_audio_per_block": (C.c_int, [C.c_void_p]),
    "fmrx_batch_reset": (C.c_int, [C.c_void_p]),
    "fmrx_batch_process": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Outputs)]),
    "fmrx_batch_process_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "fmrx_batch_sync": (C.c_int, [C.c_void_p]),
    "fmrx_batch_cuda_stream": (C.c_void_p, [C.c_void_p]),
    "fmrx_batch_cuda_stream_phase": (C.c_void_p, [C.c_void_p, C.c_int]),
    "fmrx_batch_launch_count": (C.c_longlong, [C.c_void_p]),
    "fmrx_batch_rds_offsets": (C.c_int, [C.c_void_p, i32p]),
    "fmrx_batch_tap_len": (C.c_int, [C.c_void_p, C.c_int]),
    "fmrx_batch_tap": (C.c_int, [C.c_void_p, C.c_int, fp]),
    "fmrx_batch_state_bytes": (C.c_size_t, [C.c_void_p]),
    "fmrx_batch_get_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fmrx_batch_set_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "fmrx_batch_block_id": (C.c_longlong, [C.c_void_p]),
    "fmrx_batch_partition": (C.c_int, [C.c_void_p, i32p, i32p]),
    "fmrx_batch_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Outputs), C.POINTER(C.c_longlong)]),
    "fmrx_batch_wait": (C.c_int, [C.c_void_p, C.c_longlong]),
    "fmrx_batch_wait_ingest": (C.c_int, [C.c_void_p, C.c_longlong]),
    "fmrx_batch_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "fmrx_batch_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "fmrx_batch_timeline": (C.c_int, [C.c_void_p, C.c_int, i32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "fmrx_model_firwin": (C.c_int, [C.c_int, dp, C.c_int, C.c_int, dp]),
    "fmrx_model_lfilter": (C.c_int, [dp, dp, C.c_int, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int]),
    "fmrx_model_demod": (C.c_int, [dp, dp, dp, C.c_int, C.c_int, dp]),
    "fmrx_model_pll": (C.c_int, [dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, dp]),
    "fmrx_rds_app_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "fmrx_rds_app_destroy": (None, [C.c_void_p]),
    "fmrx_rds_app_reset": (C.c_int, [C.c_void_p]),
    "fmrx_rds_app_feed": (C.c_int, [C.c_void_p, u8p, i32p, C.c_int, C.POINTER(RdsGroup), C.c_int, i32p]),
    "fmrx_rds_app_station": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(RdsStation)]),
    "fmrx_pinned_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "fmrx_pinned_free": (C.c_int, [C.c_void_p]),
    "fmrx_measure_fp32_peak": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "fmrx_measure_pll_chain": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "fmrx_ring_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fmrx_ring_destroy": (None, [C.c_void_p]),
    "fmrx_ring_acquire": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "fmrx_ring_commit": (C.c_int, [C.c_void_p]),
    "fmrx_ring_commit_blocks": (C.c_int, [C.c_void_p, C.c_int]),
    "fmrx_ring_step_blocks": (C.c_int, [C.c_void_p]),
    "fmrx_ring_close": (C.c_int, [C.c_void_p]),
    "fmrx_ring_next": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "fmrx_ring_release": (C.c_int, [C.c_void_p]),
    "fmrx_ring_in_flight": (C.c_int, [C.c_void_p]),
}

_LIB = None


def build(jobs: int = 8) -> None:
    """Compile libfmrx.so and the fm_radio CLI in-tree (nvcc, sm_100a)."""
    subprocess.check_call(["make", "-C", PKG_DIR, f"-j{jobs}", "all"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FmrxError(f"{LIB_PATH} is missing — build it with `make -C {PKG_DIR}` (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _LIB = l
    return _LIB


def check(status: int) -> None:
    if status != 0:
        raise FmrxError(f"fmrx status {status}: {lib().fmrx_last_error().decode()}")


def _p(a, t=fp):
    return a.ctypes.data_as(t)


def f32(a):
    return np.ascontiguousarray(a, dtype=F)


def _as3(a, dtype=F):
    """[n] -> [1][1][n], [B][n] -> [1][B][n], [S][B][n] unchanged."""
    a = np.ascontiguousarray(a, dtype=dtype)
    while a.ndim < 3:
        a = a[None]
    return a


# ---------------------------------------------------------------------------------------------------------------------
# thin wrappers (numpy in / numpy out; state arrays are updated in place)
# ---------------------------------------------------------------------------------------------------------------------
def design_lpf(Fs, Fc, ntaps):
    h = np.zeros(ntaps, F)
    check(lib().fmrx_design_lpf(Fs, Fc, ntaps, _p(h)))
    return h


def design_bpf(Fb, Fe, Fs, ntaps):
    h = np.zeros(ntaps, F)
    check(lib().fmrx_design_bpf(Fb, Fe, Fs, ntaps, _p(h)))
    return h


def design_rrc(Fs, ntaps):
    h = np.zeros(ntaps, F)
    check(lib().fmrx_design_rrc(Fs, ntaps, _p(h)))
    return h


def fir_response(h, Fs, f):
    """(|H(f)|, arg H(f)) of the real FIR `h` at sampling rate Fs"""
    h = f32(h)
    m, ph = C.c_double(0), C.c_double(0)
    check(lib().fmrx_fir_response(_p(h), h.size, Fs, f, C.byref(m), C.byref(ph)))
    return m.value, ph.value


def design_bpf_unity(Fb, Fe, Fs, ntaps):
    h = np.zeros(ntaps, F)
    check(lib().fmrx_design_bpf_unity(Fb, Fe, Fs, ntaps, _p(h)))
    return h


def rds_auto_phase(h_sq, Fs=240000.0, f2=114000.0):
    h = f32(h_sq)
    v = C.c_float(0)
    check(lib().fmrx_rds_auto_phase(_p(h), h.size, Fs, f2, C.byref(v)))
    return v.value


def deemphasis_coeffs(tau_us, Fs):
    b, a1 = C.c_double(0), C.c_double(0)
    check(lib().fmrx_deemphasis_coeffs(tau_us, Fs, C.byref(b), C.byref(a1)))
    return b.value, a1.value


def deemphasis(audio_f, tau_us, Fs, state, mult=1, want_int16=False):
    """audio_f: [2n] | [B][2n] | [S][B][2n] interleaved L,R; state: float32 [S][4] (x[-1], y[-1] of L then R), updated in place.
    Returns the filtered float audio (and the quantised int16 when asked)."""
    a3 = _as3(audio_f).copy()
    S, B, n2 = a3.shape
    q = np.zeros((S, B, n2), np.int16) if want_int16 else None
    check(lib().fmrx_deemphasis(_p(a3), _p(q, i16p) if want_int16 else None, S, B, n2 // 2, tau_us, Fs, mult, _p(state)))
    out = a3.reshape(np.shape(audio_f))
    return (out, q.reshape(np.shape(audio_f))) if want_int16 else out


def unpack_iq(raw):
    raw = np.ascontiguousarray(raw, np.uint8)
    out = np.zeros(raw.size, F)
    check(lib().fmrx_unpack_iq(_p(raw, u8p), raw.size, _p(out)))
    return out.reshape(raw.shape)


def fir_decim(x, h, zi, decim, exact=True):
    """x: [n] | [B][n] | [S][B][n]; zi: [nzi] | [S][nzi] (updated in place).  Returns y with x's leading shape."""
    x3, h = _as3(x), f32(h)
    S, B, n = x3.shape
    assert zi.dtype == F and zi.flags.c_contiguous and zi.size % S == 0
    y = np.zeros((S, B, n // decim), F)
    check(lib().fmrx_fir_decim(_p(y), _p(x3), S, B, n, _p(h), h.size, _p(zi), zi.size // S, decim, int(exact)))
    return y.reshape(np.shape(x)[:-1] + (n // decim,))


def fir_decim_iq(xi, xq, h, zii, ziq, decim=10, exact=True):
    a, b, h = _as3(xi), _as3(xq), f32(h)
    S, B, n = a.shape
    yi, yq = np.zeros((S, B, n // decim), F), np.zeros((S, B, n // decim), F)
    check(lib().fmrx_fir_decim_iq(_p(yi), _p(yq), _p(a), _p(b), S, B, n, _p(h), h.size, _p(zii), _p(ziq), decim, int(exact)))
    shp = np.shape(xi)[:-1] + (n // decim,)
    return yi.reshape(shp), yq.reshape(shp)


def resample(x, h, zi, decim, up, gain_up=False, ny_limit=0, exact=True):
    x3, h = _as3(x), f32(h)
    S, B, n = x3.shape
    ny = (n * up) // decim
    if 0 < ny_limit < ny:
        ny = ny_limit
    y = np.zeros((S, B, ny), F)
    check(lib().fmrx_resample(_p(y), ny_limit, _p(x3), S, B, n, _p(h), h.size, _p(zi), zi.size // S, decim, up, int(gain_up), int(exact)))
    return y.reshape(np.shape(x)[:-1] + (ny,))


def fir_mixer(nco, sig, h, zi):
    a, b, h = _as3(nco), _as3(sig), f32(h)
    S, B, n = b.shape
    y = np.zeros((S, B, n), F)
    check(lib().fmrx_fir_mixer(_p(y), _p(a), _p(b), S, B, n, _p(h), h.size, _p(zi)))
    return y.reshape(np.shape(sig))


def demod(i, q):
    a, b = _as3(i), _as3(q)
    S, B, n = a.shape
    out = np.zeros((S, B, n), F)
    check(lib().fmrx_demod(_p(a), _p(b), S, B, n, _p(out)))
    return out.reshape(np.shape(i))


def pll(x, freq, Fs, scale, phase_adj, bw, state):
    x3 = _as3(x)
    S, B, n = x3.shape
    nco = np.zeros((S, B, n), F)
    check(lib().fmrx_pll(_p(nco), _p(x3), S, B, n, freq, Fs, scale, phase_adj, bw, _p(state)))
    return nco.reshape(np.shape(x))


def pll_combine(x, h, zi, freq, Fs, scale, phase_adj, bw, state):
    x3, h = _as3(x), f32(h)
    S, B, n = x3.shape
    y, nco = np.zeros((S, B, n), F), np.zeros((S, B, n), F)
    check(lib().fmrx_pll_combine(_p(y), _p(nco), _p(x3), S, B, n, _p(h), h.size, _p(zi), freq, Fs, scale, phase_adj, bw, _p(state)))
    return y.reshape(np.shape(x)), nco.reshape(np.shape(x))


def frontend(raw, h, zii, ziq, decim=10, want_iq=False):
    """raw: u8 [2n] | [B][2n] | [S][B][2n] interleaved I,Q."""
    r3, h = _as3(raw, np.uint8), f32(h)
    S, B, n2 = r3.shape
    n = n2 // 2
    d = np.zeros((S, B, n // decim), F)
    yi = np.zeros_like(d) if want_iq else None
    yq = np.zeros_like(d) if want_iq else None
    check(lib().fmrx_frontend(_p(d), _p(yi) if want_iq else None, _p(yq) if want_iq else None, _p(r3, u8p), S, B, n, _p(h), h.size, _p(zii), _p(ziq), decim))
    shp = np.shape(raw)[:-1] + (n // decim,)
    return (d.reshape(shp), yi.reshape(shp), yq.reshape(shp)) if want_iq else d.reshape(shp)


def rds_decode(rrc, state):
    """rrc: [n] | [B][n] | [S][B][n]; state: int32 [S][160] updated in place.
    Returns (bits [S][B][80] u8, n_bits [S][B], events [S][B][MAX_EVENTS] structured, n_events [S][B])."""
    r3 = _as3(rrc)
    S, B, n = r3.shape
    bits = np.zeros((S, B, MAX_BITS), np.uint8)
    nb = np.zeros((S, B), np.int32)
    ev = np.zeros((S, B, MAX_EVENTS), EVENT_DTYPE)
    ne = np.zeros((S, B), np.int32)
    check(lib().fmrx_rds_decode(_p(r3), S, B, n, _p(bits, u8p), _p(nb, i32p), _p(ev, evp), _p(ne, i32p), _p(state, i32p)))
    return bits, nb, ev, ne


def rds_format_block(block_id, initial_offset, events):
    ev = np.ascontiguousarray(events, EVENT_DTYPE)
    buf = C.create_string_buffer(16384)
    n = lib().fmrx_rds_format_block(block_id, initial_offset, _p(ev, evp) if ev.size else None, ev.size, buf, 16384)
    return buf.raw[:n].decode()


def measure_pll_chain(device=0):
    """SM cycles per step of one PLL loop's dependency chain, run alone from registers (the PLL kernel's latency roofline)."""
    v = C.c_double(0)
    check(lib().fmrx_measure_pll_chain(device, C.byref(v)))
    return v.value


def measure_fp32_peak(kind, device=0, reps=5):
    v = C.c_double(0)
    check(lib().fmrx_measure_fp32_peak(device, kind, reps, C.byref(v)))
    return v.value


# ---------------------------------------------------------------------------------------------------------------------
# the batched chain
# ---------------------------------------------------------------------------------------------------------------------
# ---- model-compatible operators (float64, the arithmetic of the reference's Python models)
def _d(a):
    return np.ascontiguousarray(a, np.float64)


def model_firwin(ntaps, cutoff, pass_zero=True):
    """scipy.signal.firwin(ntaps, cutoff, window='hann', pass_zero=...) with cutoff a scalar (low-pass) or [lo, hi] (band-pass)"""
    c = _d(np.atleast_1d(cutoff))
    h = np.empty(ntaps, np.float64)
    check(lib().fmrx_model_firwin(ntaps, c.ctypes.data_as(dp), c.size, 1 if pass_zero is True or pass_zero == "lowpass" else 0, h.ctypes.data_as(dp)))
    return h


def model_lfilter(x, b, hist, decim=1, up=1):
    """lfilter(b, 1, x, zi) then [::decim] (on the input zero-stuffed by `up`); x:[S][n] or [n]; hist:[S][len(b)-1] updated in place"""
    x2 = _d(x)
    x2 = x2.reshape(1, -1) if x2.ndim == 1 else x2
    b = _d(b)
    S, n = x2.shape
    assert hist.dtype == np.float64 and hist.flags.c_contiguous and hist.size == S * (b.size - 1)
    y = np.empty((S, n * up // decim), np.float64)
    check(lib().fmrx_model_lfilter(y.ctypes.data_as(dp), x2.ctypes.data_as(dp), S, n, b.ctypes.data_as(dp), b.size, hist.ctypes.data_as(dp), decim, up))
    return y[0] if np.ndim(x) == 1 else y


def model_demod(i, q, prev_phase):
    """fmSupportLib.fmDemodArctan; prev_phase: float64 array [S] updated in place"""
    i2, q2 = _d(i), _d(q)
    i2 = i2.reshape(1, -1) if i2.ndim == 1 else i2
    q2 = q2.reshape(i2.shape)
    assert prev_phase.dtype == np.float64 and prev_phase.size == i2.shape[0]
    out = np.empty_like(i2)
    check(lib().fmrx_model_demod(out.ctypes.data_as(dp), i2.ctypes.data_as(dp), q2.ctypes.data_as(dp), i2.shape[0], i2.shape[1], prev_phase.ctypes.data_as(dp)))
    return out[0] if np.ndim(i) == 1 else out


def model_pll(x, freq, Fs, state, nco_scale=1.0, phase_adjust=0.0, norm_bandwidth=0.01):
    """fmPll.fmPll; state: float64 [S][6] in the model's order, updated in place; returns (ncoOut, ncoOutQ) of length n+1"""
    x2 = _d(x)
    x2 = x2.reshape(1, -1) if x2.ndim == 1 else x2
    S, n = x2.shape
    assert state.dtype == np.float64 and state.size == 6 * S
    a, b = np.empty((S, n + 1), np.float64), np.empty((S, n + 1), np.float64)
    check(lib().fmrx_model_pll(a.ctypes.data_as(dp), b.ctypes.data_as(dp), x2.ctypes.data_as(dp), S, n, freq, Fs, nco_scale, phase_adjust, norm_bandwidth,
                               state.ctypes.data_as(dp)))
    return (a[0], b[0]) if np.ndim(x) == 1 else (a, b)


class RdsApp:
    """RDS data-link / application layer over the bits a Batch returns (host code, fmrx_rds_app_*)."""

    def __init__(self, n_streams=1):
        self.S, self.h = n_streams, C.c_void_p()
        check(lib().fmrx_rds_app_create(n_streams, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().fmrx_rds_app_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        check(lib().fmrx_rds_app_reset(self.h))

    def feed(self, bits, n_bits, cap=64):
        """bits [S][B][MAX_BITS] u8, n_bits [S][B] i32 (as in Batch.process results) -> list (per stream) of group dicts"""
        bits = np.ascontiguousarray(bits, np.uint8).reshape(self.S, -1, MAX_BITS)
        n_bits = np.ascontiguousarray(n_bits, np.int32).reshape(self.S, -1)
        g = (RdsGroup * (self.S * cap))()
        ng = np.zeros(self.S, np.int32)
        check(lib().fmrx_rds_app_feed(self.h, bits.ctypes.data_as(u8p), n_bits.ctypes.data_as(i32p), bits.shape[1], g, cap, ng.ctypes.data_as(i32p)))
        return [[dict(blk=list(g[s * cap + i].blk), type=g[s * cap + i].type, version_b=g[s * cap + i].version_b, corrected=g[s * cap + i].corrected,
                      bit_index=g[s * cap + i].bit_index) for i in range(min(int(ng[s]), cap))] for s in range(self.S)]

    def feed_stream(self, stream_bits, cap=4096):
        """convenience for one station: a flat bit array, cut into MAX_BITS-sized pieces"""
        assert self.S == 1
        b = np.asarray(stream_bits, np.uint8).ravel()
        nblk = max(1, -(-b.size // MAX_BITS))
        pad = np.zeros(nblk * MAX_BITS, np.uint8)
        pad[:b.size] = b
        nb = np.full(nblk, MAX_BITS, np.int32)
        if b.size % MAX_BITS or b.size == 0:
            nb[-1] = b.size - (nblk - 1) * MAX_BITS
        return self.feed(pad.reshape(1, nblk, MAX_BITS), nb.reshape(1, nblk), cap)[0]

    def station(self, stream=0):
        st = RdsStation()
        check(lib().fmrx_rds_app_station(self.h, stream, C.byref(st)))
        return dict(synced=bool(st.synced), pi=st.pi, pty=st.pty, tp=st.tp, ps=st.ps.decode("latin-1"), rt=st.rt.decode("latin-1"), ps_complete=bool(st.ps_complete),
                    groups=st.groups, blocks_ok=st.blocks_ok, blocks_corrected=st.blocks_corrected, blocks_bad=st.blocks_bad, sync_losses=st.sync_losses,
                    bits_fed=st.bits_fed)


class Batch:
    """n_streams independent stations processed in lock step, max_blocks blocks per call."""

    def __init__(self, n_streams=1, mode=0, profile=PROFILE_BINARY, max_blocks=1, device=0, paths=0, numerics=NUMERICS_REFERENCE, quality=0):
        self.cfg = Config(mode, profile, n_streams, max_blocks, device, paths, numerics, quality)
        self.h = C.c_void_p()
        check(lib().fmrx_batch_create(C.byref(self.cfg), C.byref(self.h)))
        self.S, self.mode, self.max_blocks = n_streams, mode, max_blocks
        self.n_audio = lib().fmrx_batch_audio_per_block(self.h)
        self.rds = mode != 1 and (paths == 0 or paths & PATH_RDS)
        self.audio_on = paths == 0 or bool(paths & PATH_AUDIO)
        self.block_id = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            lib().fmrx_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown: module globals may already be gone
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        check(lib().fmrx_batch_reset(self.h))
        self.block_id = 0

    def process(self, iq, want_float=False):
        """iq: u8 [S][B*307200] (or [B*307200] for one stream).  Returns a dict with audio int16 [S][B][2*n_audio]
        (L,R interleaved), optionally audio_f, and for mode 0 rds_bits / rds_n_bits / rds_events / rds_n_events."""
        iq = np.ascontiguousarray(iq, np.uint8).reshape(self.S, -1)
        assert iq.shape[1] % BLOCK_BYTES == 0, "whole 307200-byte blocks only (SURVEY Q9)"
        B = iq.shape[1] // BLOCK_BYTES
        res, o = {}, Outputs()
        if self.audio_on:
            res["audio"] = np.zeros((self.S, B, 2 * self.n_audio), np.int16)
            o.audio = _p(res["audio"], i16p)
            if want_float:
                res["audio_f"] = np.zeros((self.S, B, 2 * self.n_audio), F)
                o.audio_f = _p(res["audio_f"])
        if self.rds:
            res["rds_bits"] = np.zeros((self.S, B, MAX_BITS), np.uint8); o.rds_bits = _p(res["rds_bits"], u8p)
            res["rds_n_bits"] = np.zeros((self.S, B), np.int32); o.rds_n_bits = _p(res["rds_n_bits"], i32p)
            res["rds_events"] = np.zeros((self.S, B, MAX_EVENTS), EVENT_DTYPE); o.rds_events = _p(res["rds_events"], evp)
            res["rds_n_events"] = np.zeros((self.S, B), np.int32); o.rds_n_events = _p(res["rds_n_events"], i32p)
        check(lib().fmrx_batch_process(self.h, iq.ctypes.data_as(C.c_void_p), B, C.byref(o)))
        res["first_block"] = self.block_id
        self.block_id += B
        self.last_blocks = B
        return res

    def process_device(self, iq_ptr: int, n_blocks: int, out_ptrs: Outputs | None = None):
        """Raw device pointers (e.g. torch tensor .data_ptr()); asynchronous — call sync()."""
        check(lib().fmrx_batch_process_device(self.h, C.c_void_p(iq_ptr), n_blocks, C.byref(out_ptrs) if out_ptrs is not None else None))
        self.block_id += n_blocks
        self.last_blocks = n_blocks

    def sync(self):
        check(lib().fmrx_batch_sync(self.h))

    def tap(self, name):
        which = TAPS[name]
        n = lib().fmrx_batch_tap_len(self.h, which)
        out = np.zeros((self.S, self.last_blocks, n), F)
        check(lib().fmrx_batch_tap(self.h, which, _p(out)))
        return out

    def rds_offsets(self):
        out = np.zeros(self.S, np.int32)
        check(lib().fmrx_batch_rds_offsets(self.h, _p(out, i32p)))
        return out

    @property
    def rds_phase(self):
        return lib().fmrx_batch_rds_phase(self.h)

    @rds_phase.setter
    def rds_phase(self, v):
        check(lib().fmrx_batch_set_rds_phase(self.h, float(v)))

    def rds_text(self, res, stream=0):
        """The stderr lines the reference's frame_thread prints for the blocks of `res` (src/fm_radio.cpp:516,619-701)."""
        off = int(self.rds_offsets()[stream])
        out = []
        for b in range(res["rds_n_events"].shape[1]):
            ne = int(res["rds_n_events"][stream, b])
            out.append(rds_format_block(res["first_block"] + b, off, res["rds_events"][stream, b, :ne]))
        return "".join(out)

    def partition(self):
        """(SMs owned by the PLL phase, SMs owned by the filter phases); (0, 0) when the phases share the device."""
        a, b = C.c_int32(), C.c_int32()
        check(lib().fmrx_batch_partition(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile(self, enable=True):
        """True / 1: per-stage timing with the phases serialised; 2: timeline mode (pipeline kept); False: off."""
        check(lib().fmrx_batch_profile(self.h, int(enable)))

    def timeline(self, cap=4096):
        """[(stage name, start ms, end ms)] of every bracket since profile(2) (pipelined timeline mode)."""
        st = (C.c_int32 * cap)()
        t0 = (C.c_float * cap)()
        t1 = (C.c_float * cap)()
        n = lib().fmrx_batch_timeline(self.h, cap, st, t0, t1)
        if n < 0:
            check(n)
        return [(STAGES[st[i]], t0[i], t1[i]) for i in range(n)]

    def stage_times(self):
        """{stage: (total ms, brackets)} accumulated since profile(True)."""
        ms = (C.c_double * len(STAGES))()
        cnt = (C.c_longlong * len(STAGES))()
        check(lib().fmrx_batch_stage_times(self.h, ms, cnt))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(STAGES)}

    def process_into(self, iq, n_blocks, out: "Outputs"):
        """Host path with caller-provided (ideally pinned) buffers; `iq` is anything with a ctypes-able address."""
        ptr = iq if isinstance(iq, int) else iq.ctypes.data
        check(lib().fmrx_batch_process(self.h, C.c_void_p(ptr), n_blocks, C.byref(out)))
        self.block_id += n_blocks
        self.last_blocks = n_blocks

    def submit(self, iq, n_blocks, out: "Outputs"):
        """Asynchronous host path: enqueue one step (pinned buffers), return its ticket; wait(ticket) before reading `out`."""
        ptr = iq if isinstance(iq, int) else iq.ctypes.data
        t = C.c_longlong(-1)
        check(lib().fmrx_batch_submit(self.h, C.c_void_p(ptr), n_blocks, C.byref(out), C.byref(t)))
        self.block_id += n_blocks
        self.last_blocks = n_blocks
        return t.value

    def wait(self, ticket):
        check(lib().fmrx_batch_wait(self.h, ticket))

    def wait_ingest(self, ticket):
        """returns when that step's host-to-device copy has finished (its input buffer may be refilled)"""
        check(lib().fmrx_batch_wait_ingest(self.h, ticket))

    @property
    def launches(self):
        return lib().fmrx_batch_launch_count(self.h)

    def get_state(self):
        buf = np.zeros(lib().fmrx_batch_state_bytes(self.h), np.uint8)
        check(lib().fmrx_batch_get_state(self.h, buf.ctypes.data_as(C.c_void_p), buf.size))
        return buf

    def set_state(self, blob, block_id=None):
        blob = np.ascontiguousarray(blob, np.uint8)
        check(lib().fmrx_batch_set_state(self.h, blob.ctypes.data_as(C.c_void_p), blob.size))
        self.block_id = int(lib().fmrx_batch_block_id(self.h))


ERR_TIMEOUT, ERR_EOF = -5, -6


class Ring:
    """Bounded ring of pinned host slots in front of a Batch (fmrx_ring_*): one producer thread acquires / fills /
    commits, one consumer thread takes the steps in order.  `acquire` and `next` return numpy views of the slot's
    pinned buffers; `next` returns None at end of input (after close()).  timeout_ms < 0 waits for ever; a timeout
    raises TimeoutError."""

    def __init__(self, batch: "Batch", n_slots=4, n_blocks=1):
        self.batch, self.n_blocks = batch, n_blocks
        self.h = C.c_void_p()
        check(lib().fmrx_ring_create(batch.h, n_slots, n_blocks, C.byref(self.h)))

    def close_ring(self):
        if getattr(self, "h", None) and self.h.value:
            lib().fmrx_ring_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close_ring()

    @staticmethod
    def _status(st):
        if st == ERR_TIMEOUT:
            raise TimeoutError(lib().fmrx_last_error().decode())
        check(st)

    def acquire(self, timeout_ms=-1):
        p = C.c_void_p()
        self._status(lib().fmrx_ring_acquire(self.h, timeout_ms, C.byref(p)))
        n = self.batch.S * self.n_blocks * BLOCK_BYTES
        return np.ctypeslib.as_array(C.cast(p, u8p), shape=(n,)).reshape(self.batch.S, self.n_blocks * BLOCK_BYTES)

    def commit(self, n_blocks=None):
        check(lib().fmrx_ring_commit(self.h) if n_blocks is None else lib().fmrx_ring_commit_blocks(self.h, n_blocks))

    def close(self):
        check(lib().fmrx_ring_close(self.h))

    def next(self, timeout_ms=-1):
        o = Outputs()
        st = lib().fmrx_ring_next(self.h, timeout_ms, C.byref(o))
        if st == ERR_EOF:
            return None
        self._status(st)
        S, B, na = self.batch.S, lib().fmrx_ring_step_blocks(self.h), self.batch.n_audio
        res = {}
        if o.audio:
            res["audio"] = np.ctypeslib.as_array(o.audio, shape=(S, B, 2 * na))
        if o.rds_bits:
            res["rds_bits"] = np.ctypeslib.as_array(o.rds_bits, shape=(S, B, MAX_BITS))
            res["rds_n_bits"] = np.ctypeslib.as_array(o.rds_n_bits, shape=(S, B))
            res["rds_n_events"] = np.ctypeslib.as_array(o.rds_n_events, shape=(S, B))
            res["rds_events"] = np.ctypeslib.as_array(C.cast(o.rds_events, i32p), shape=(S, B, MAX_EVENTS, 4)).view(EVENT_DTYPE)[..., 0]
        return res

    def release(self):
        check(lib().fmrx_ring_release(self.h))

    @property
    def in_flight(self):
        return lib().fmrx_ring_in_flight(self.h)
