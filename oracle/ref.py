"""TEST INFRASTRUCTURE ONLY — ctypes binding of the UNMODIFIED reference objects behind oracle/ref_shim.cpp
(oracle/_ref/libfmref.so) and a runner for the reference binary (oracle/_ref/fm_radio).

`RefChain` composes the reference FUNCTIONS exactly as the thread bodies of src/fm_radio.cpp do (call order, argument
values and buffer lifetimes cited inline), in the two profiles of SURVEY App. A:
  binary — what the shipped executable observably does (stereo dead from block 1 on, Q7);
  intent — same functions, caller-side UB repaired (mixed stays sized, outputs assigned not accumulated).
It exists to pin the oracle port where the binary itself gives no observable output (stereo after block 0, every
intermediate signal, the RDS float stages).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libfmref.so")
REF_BIN = os.path.join(HERE, "_ref", "fm_radio")
F = np.float32
fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
PI = 3.14159265358979323846


def ref_available() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


def ref_binary() -> str:
    return REF_BIN


def run_ref_binary(raw: np.ndarray, mode: int = 0, timeout: float = 600.0):
    """Feeds `raw` to the reference executable.  Mode 0 = no argument, mode 1 = "1" (src/fm_radio.cpp:736-764)."""
    args = [REF_BIN] + (["1"] if mode == 1 else [])
    r = subprocess.run(args, input=np.ascontiguousarray(raw, np.uint8).tobytes(), capture_output=True, timeout=timeout)
    return np.frombuffer(r.stdout, dtype=np.int16).copy(), r.stderr.decode(), r.returncode


def _p(a, t=fp):
    return a.ctypes.data_as(t)


def f32(a):
    return np.ascontiguousarray(a, dtype=F)


_LIB = None


def load_ref():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(REF_SO)
        lib.ref_lpf.argtypes = [C.c_float, C.c_float, C.c_ushort, fp]
        lib.ref_bpf.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, fp]
        lib.ref_rrc.argtypes = [C.c_float, C.c_int, fp]
        lib.ref_fmpll.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        lib.ref_pll_combine.argtypes = [fp, fp, fp, C.c_int, fp, C.c_int, fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        lib.ref_conv_mixer.argtypes = [fp, fp, fp, C.c_int, C.c_int, C.c_float, fp, C.c_int, fp, C.c_int]
        _LIB = lib
    return _LIB


class Ref:
    def __init__(self):
        self.lib = load_ref()

    def lpf(self, Fs, Fc, ntaps):
        h = np.zeros(ntaps, F); self.lib.ref_lpf(Fs, Fc, ntaps, _p(h)); return h

    def bpf(self, Fb, Fe, Fs, ntaps):
        h = np.zeros(ntaps, F); self.lib.ref_bpf(Fb, Fe, Fs, ntaps, _p(h)); return h

    def rrc(self, Fs, ntaps):
        h = np.zeros(ntaps, F); self.lib.ref_rrc(Fs, ntaps, _p(h)); return h

    def unpack(self, raw, nsamples=None):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        nsamples = raw.size if nsamples is None else nsamples
        out = np.zeros(nsamples, F)
        self.lib.ref_unpack(_p(raw, u8p), raw.size, nsamples, _p(out))
        return out

    def fir_decim(self, x, h, zi, decim, y_init=None):
        x, h = f32(x), f32(h)
        y = np.zeros(x.size // decim, F)
        yi = _p(f32(y_init)) if y_init is not None else None
        self.lib.ref_conv_decim(yi, _p(y), _p(x), x.size, _p(h), h.size, _p(zi), decim)
        return y

    def fir_decim_ptr(self, x, h, zi, decim, y_init=None):
        x, h = f32(x), f32(h)
        y = np.zeros(x.size // decim, F)
        yi = _p(f32(y_init)) if y_init is not None else None
        self.lib.ref_conv_decim_ptr(yi, _p(y), _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim)
        return y

    def fir_decim_iq(self, xi, xq, h, zii, ziq, decim):
        xi, xq, h = f32(xi), f32(xq), f32(h)
        yi = np.zeros(xi.size // decim, F); yq = np.zeros_like(yi)
        self.lib.ref_conv_decim_iq(_p(yi), _p(yq), _p(xi), _p(xq), xi.size, _p(h), h.size, _p(zii), _p(ziq), decim)
        return yi, yq

    def resample(self, x, h, zi, decim, up, ny_keep=None):
        x, h = f32(x), f32(h)
        ny = (x.size * up) // decim
        keep = ny if ny_keep is None else ny_keep
        y = np.zeros(keep, F)
        self.lib.ref_conv_mode1(_p(y), keep, _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim, up)
        return y

    def resample_ptr(self, x, h, zi, decim, up):
        x, h = f32(x), f32(h)
        y = np.zeros((x.size * up) // decim, F)
        self.lib.ref_conv_mode1_ptr(_p(y), _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim, up)
        return y

    def resample_rds(self, x, h, zi, decim, up):
        x, h = f32(x), f32(h)
        y = np.zeros((x.size * up) // decim, F)
        self.lib.ref_conv_mode1_rds(_p(y), _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim, up)
        return y

    def fir_mixer(self, nco, sig, h, zi, pad=0.0):
        """nco has len(sig)+1 elements (untrimmed pllCombine output); returns len(sig)+1 outputs like the reference."""
        nco, sig, h = f32(nco), f32(sig), f32(h)
        y = np.zeros(nco.size, F)
        self.lib.ref_conv_mixer(_p(y), _p(nco), _p(sig), nco.size, sig.size, pad, _p(h), h.size, _p(zi), 1)
        return y

    def demod(self, i, q):
        i, q = f32(i), f32(q)
        out = np.zeros(i.size, F)
        self.lib.ref_demod(_p(i), _p(q), i.size, _p(out))
        return out

    def pll(self, x, freq, Fs, scale, phase_adj, bw, st):
        x = f32(x)
        nco = np.zeros(x.size, F)
        self.lib.ref_fmpll(_p(nco), _p(x), x.size, freq, Fs, scale, phase_adj, bw, _p(st))
        return nco

    def pll_combine(self, x, h, zi, freq, Fs, scale, phase_adj, bw, st):
        x, h = f32(x), f32(h)
        y = np.zeros(x.size, F); nco = np.zeros(x.size + 1, F)
        self.lib.ref_pll_combine(_p(y), _p(nco), _p(x), x.size, _p(h), h.size, _p(zi), 1, freq, Fs, scale, phase_adj, bw, _p(st))
        return y, nco

    def frame_thread(self, rrc_blocks):
        """Runs the real frame_thread body (src/fm_radio.cpp:444-729) over the given RRC blocks; returns its stderr."""
        rrc = f32(rrc_blocks)
        nblk, blk_len = rrc.shape
        buf = C.create_string_buffer(1 << 20)
        n = self.lib.ref_frame_thread(_p(rrc), nblk, blk_len, buf, 1 << 20)
        return buf.raw[:min(n, (1 << 20) - 1)].decode()


class RefChain:
    NT = 151
    NIF = 15360

    def __init__(self, mode=0, profile=0):
        r = self.r = Ref()
        self.mode, self.profile, self.block_id = mode, profile, 0
        nt = self.NT
        self.rf_fs = 2500000 if mode == 1 else 2400000            # src/fm_radio.cpp:36-37
        self.zi_i, self.zi_q = np.zeros(nt - 1, F), np.zeros(nt - 1, F)
        audio_fs, self.audio_taps, self.decim, self.up, self.mult = 240000, nt, 5, 1, 1   # :153-162
        if mode == 1:
            audio_fs, self.decim, self.up, self.audio_taps, self.mult = 6000000, 125, 24, nt * 24, 24   # :174-180, :229
        self.h_mono = r.lpf(audio_fs, 16000, self.audio_taps)     # :200
        self.h_pilot = r.bpf(18.5e3, 19.5e3, audio_fs, nt)        # :201
        self.h_sbpf = r.bpf(22e3, 54e3, audio_fs, nt)             # :202
        self.h_stereo = r.lpf(audio_fs, 16000, self.audio_taps)   # :203
        z = self.audio_taps - 1                                   # :189-193
        self.zi_mono, self.zi_pilot, self.zi_sbpf, self.zi_stereo = (np.zeros(z, F) for _ in range(4))
        self.pll_st = np.array([0, 0, 1, 0, 0, 1], F)             # :165-171
        # rds_thread :331-370
        self.h_rbpf = r.bpf(54000, 60000, 240000, nt)
        self.h_sq = r.bpf(113500, 114500, 240000, nt)
        self.h_lpf = r.lpf(240000, 3000, nt)
        self.h_anti = r.lpf(float(F(240000) * F(19)), 57000 // 2, nt * 19)
        self.h_rrc = r.rrc(57000, nt)
        self.zi_rbpf, self.zi_sq, self.zi_lpf, self.zi_rrc = (np.zeros(nt - 1, F) for _ in range(4))
        self.zi_anti = np.zeros(nt * 19 - 1, F)
        self.rds_pll_st = np.array([0, 0, 1, 0, 0, 1], F)         # :343-349
        phase_adj = F(PI / 3.3 - PI / 1.5)                        # :342
        self.rds_phase = float(F(float(phase_adj) - PI / 1.4))    # :400
        self.taps = {}

    def block(self, raw, rds=True):
        r, t = self.r, self.taps
        raw = np.ascontiguousarray(raw, np.uint8)
        iq = r.unpack(raw)                                        # :66
        i_data, q_data = iq[0::2].copy(), iq[1::2].copy()         # :68-72
        h_rf = r.lpf(self.rf_fs, 100000, self.NT)                 # :75 (re-designed every block)
        t["i"], t["q"] = r.fir_decim_iq(i_data, q_data, h_rf, self.zi_i, self.zi_q, 10)   # :78
        demod = t["demod"] = r.demod(t["i"], t["q"])              # :84
        # ---- mono_stero_thread
        if self.mode == 1:
            mono = r.resample_ptr(demod, self.h_mono, self.zi_mono, self.decim, self.up)      # :228
        else:
            mono = r.fir_decim_ptr(demod, self.h_mono, self.zi_mono, 5)                       # :258
        t["mono"] = mono
        na = mono.size
        if self.profile == 1 or self.block_id == 0:
            t["pilot"] = r.fir_decim_ptr(demod, self.h_pilot, self.zi_pilot, 1)               # :232 / :261
            t["nco"] = r.pll(t["pilot"], 19e3, 240e3, 2.0, 0.0, 0.01, self.pll_st)            # :233 / :262
            t["stereo_bpf"] = r.fir_decim_ptr(demod, self.h_sbpf, self.zi_sbpf, 1)            # :236 / :265
            mixed = (t["stereo_bpf"] * t["nco"]).astype(F)                                    # :240-243 / :269-272
            if self.mode == 1:
                st = r.resample(mixed, self.h_stereo, self.zi_stereo, 5, self.up, ny_keep=na)  # :245
            else:
                st = r.fir_decim(mixed, self.h_stereo, self.zi_stereo, 5)                     # :274
        else:
            st = np.zeros(na, F)                                                              # Q7
        t["stereo"] = st
        left = ((mono + st) / F(2)).astype(F)                                                 # :250 / :280
        right = ((mono - st) / F(2)).astype(F)                                                # :251 / :281
        lr = np.empty(2 * na, F); lr[0::2] = left; lr[1::2] = right
        t["audio_f"] = lr
        with np.errstate(invalid="ignore", over="ignore"):
            scaled = (lr * F(16384)) * F(self.mult)                                           # :297
            q = np.where(np.isnan(lr), 0, np.trunc(np.nan_to_num(scaled))).astype(np.int64)   # :290-298
        audio = (q & 0xFFFF).astype(np.uint16).view(np.int16)
        # ---- rds_thread
        if self.mode == 0 and rds:
            t["rds_bpf"] = r.fir_decim_ptr(demod, self.h_rbpf, self.zi_rbpf, 1)               # :395
            t["rds_sq"], t["rds_nco"] = r.pll_combine(t["rds_bpf"], self.h_sq, self.zi_sq, 114000, 240000, 0.5,
                                                      self.rds_phase, 0.001, self.rds_pll_st)  # :400
            lp = r.fir_mixer(t["rds_nco"], t["rds_bpf"], self.h_lpf, self.zi_lpf)             # :404 (15361 long)
            t["rds_lpf"] = lp[:-1]
            t["rds_res"] = r.resample_rds(lp, self.h_anti, self.zi_anti, 80, 19)              # :408
            t["rds_rrc"] = r.fir_decim(t["rds_res"], self.h_rrc, self.zi_rrc, 1)              # :411
        self.block_id += 1
        return audio
