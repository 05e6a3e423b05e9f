"""TEST INFRASTRUCTURE ONLY — ctypes binding of the oracle port (fmrx_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F = np.float32
fp = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)

TAPS = dict(demod=0, mono=1, pilot=2, nco=3, stereo_bpf=4, stereo=5, rds_bpf=6, rds_sq=7, rds_nco=8, rds_lpf=9, rds_res=10,
            rds_rrc=11, i=12, q=13, audio_f=14)


class Event(C.Structure):
    _fields_ = [("block", C.c_int32), ("kind", C.c_int32), ("letter", C.c_int32), ("position", C.c_uint32)]

    def astuple(self):
        return (self.block, self.kind, self.letter, self.position)


def build(force: bool = False) -> None:
    """Compile the port (and, if /root/reference is mounted, oracle/_ref) via oracle/Makefile."""
    so = os.path.join(HERE, "libfmrx_oracle.so")
    src = os.path.join(HERE, "fmrx_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/src/fm_radio.cpp"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


def _p(a, t=fp):
    return a.ctypes.data_as(t)


def f32(a):
    return np.ascontiguousarray(a, dtype=F)


_LIB = None


def load_port():
    global _LIB
    if _LIB is None:
        build()
        lib = C.CDLL(os.path.join(HERE, "libfmrx_oracle.so"))
        lib.orc_chain_create.restype = C.c_void_p
        lib.orc_chain_tap.restype = fp
        lib.orc_rds_decoder_create.restype = C.c_void_p
        lib.orc_design_lpf.argtypes = [C.c_float, C.c_float, C.c_ushort, fp]
        lib.orc_design_bpf.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, fp]
        lib.orc_design_rrc.argtypes = [C.c_float, C.c_int, fp]
        lib.orc_pll.argtypes = [fp, fp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        lib.orc_pll_combine.argtypes = [fp, fp, fp, C.c_int, fp, C.c_int, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, fp]
        _LIB = lib
    return _LIB


class Port:
    """Function-level view of the port; every method mirrors one reference function (see fmrx_oracle.h)."""

    def __init__(self):
        self.lib = load_port()

    def lpf(self, Fs, Fc, ntaps):
        h = np.zeros(ntaps, F)
        self.lib.orc_design_lpf(Fs, Fc, ntaps, _p(h))
        return h

    def bpf(self, Fb, Fe, Fs, ntaps):
        h = np.zeros(ntaps, F)
        self.lib.orc_design_bpf(Fb, Fe, Fs, ntaps, _p(h))
        return h

    def rrc(self, Fs, ntaps):
        h = np.zeros(ntaps, F)
        self.lib.orc_design_rrc(Fs, ntaps, _p(h))
        return h

    def unpack(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        out = np.zeros(raw.size, F)
        self.lib.orc_unpack(_p(raw, u8p), C.c_size_t(raw.size), _p(out))
        return out

    def fir_decim(self, x, h, zi, decim):
        x, h = f32(x), f32(h)
        y = np.zeros(x.size // decim, F)
        self.lib.orc_fir_decim(_p(y), _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim)
        return y

    def fir_decim_iq(self, xi, xq, h, zii, ziq, decim):
        xi, xq, h = f32(xi), f32(xq), f32(h)
        yi = np.zeros(xi.size // decim, F)
        yq = np.zeros_like(yi)
        self.lib.orc_fir_decim_iq(_p(yi), _p(yq), _p(xi), _p(xq), xi.size, _p(h), h.size, _p(zii), _p(ziq), decim)
        return yi, yq

    def resample(self, x, h, zi, decim, up, gain_up=False, ny_limit=0):
        x, h = f32(x), f32(h)
        ny = (x.size * up) // decim
        if 0 < ny_limit < ny:
            ny = ny_limit
        y = np.zeros(ny, F)
        self.lib.orc_resample(_p(y), ny_limit, _p(x), x.size, _p(h), h.size, _p(zi), zi.size, decim, up, int(gain_up))
        return y

    def fir_mixer(self, nco, sig, h, zi):
        nco, sig, h = f32(nco), f32(sig), f32(h)
        y = np.zeros(sig.size, F)
        self.lib.orc_fir_mixer(_p(y), _p(nco), _p(sig), sig.size, _p(h), h.size, _p(zi))
        return y

    def demod(self, i, q):
        i, q = f32(i), f32(q)
        out = np.zeros(i.size, F)
        self.lib.orc_demod(_p(i), _p(q), i.size, _p(out))
        return out

    def pll(self, x, freq, Fs, scale, phase_adj, bw, st):
        x = f32(x)
        nco = np.zeros(x.size, F)
        self.lib.orc_pll(_p(nco), _p(x), x.size, freq, Fs, scale, phase_adj, bw, _p(st))
        return nco

    def pll_combine(self, x, h, zi, freq, Fs, scale, phase_adj, bw, st):
        x, h = f32(x), f32(h)
        y = np.zeros(x.size, F)
        nco = np.zeros(x.size + 1, F)
        self.lib.orc_pll_combine(_p(y), _p(nco), _p(x), x.size, _p(h), h.size, _p(zi), freq, Fs, scale, phase_adj, bw, _p(st))
        return y, nco


def format_block(block_id, initial_offset, events):
    lib = load_port()
    ev = (Event * max(1, len(events)))(*[Event(*e) for e in events])
    buf = C.create_string_buffer(16384)
    n = lib.orc_rds_format_block(block_id, initial_offset, ev, len(events), buf, 16384)
    return buf.raw[:n].decode()


class RdsDecoder:
    def __init__(self):
        self.lib = load_port()
        self.h = C.c_void_p(self.lib.orc_rds_decoder_create())
        self.block = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_rds_decoder_destroy(self.h)
            self.h = None

    def block_decode(self, rrc):
        rrc = f32(rrc)
        bits = np.zeros(256, np.uint8)
        ev = (Event * 256)()
        nev = C.c_int(0)
        nb = self.lib.orc_rds_decode_block(self.h, _p(rrc), rrc.size, _p(bits, u8p), ev, 256, C.byref(nev))
        self.block += 1
        return bits[:nb].copy(), [ev[i].astuple() for i in range(nev.value)]

    @property
    def initial_offset(self):
        return self.lib.orc_rds_initial_offset(self.h)

    @property
    def start_pos(self):
        return self.lib.orc_rds_start_pos(self.h)


class Chain:
    """Sequential composition of the four thread bodies of src/fm_radio.cpp (profile 0 = binary, 1 = intent)."""

    def __init__(self, mode=0, profile=0, paths=3, quality=0):
        self.lib = load_port()
        self.h = C.c_void_p(self.lib.orc_chain_create(mode, profile))
        self.lib.orc_chain_set_paths(self.h, paths)
        if quality:
            self.lib.orc_chain_set_quality(self.h, quality)
        self.mode = mode
        self.n_audio = self.lib.orc_chain_audio_per_block(self.h)
        self.block_id = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_chain_destroy(self.h)
            self.h = None

    def block(self, raw):
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        assert raw.size == 307200
        audio = np.zeros(2 * self.n_audio, np.int16)
        self.lib.orc_chain_block(self.h, _p(raw, u8p), audio.ctypes.data_as(C.POINTER(C.c_int16)))
        self.block_id += 1
        return audio

    def tap(self, name):
        n = C.c_int(0)
        p = self.lib.orc_chain_tap(self.h, TAPS[name], C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def rds(self):
        bits = np.zeros(256, np.uint8)
        nb = self.lib.orc_chain_rds_bits(self.h, _p(bits, u8p), 256)
        ev = (Event * 256)()
        ne = self.lib.orc_chain_rds_events(self.h, ev, 256)
        return bits[:nb].copy(), [ev[i].astuple() for i in range(ne)]

    @property
    def rds_offset(self):
        return self.lib.orc_chain_rds_offset(self.h)

    def run(self, raw, taps=()):
        """Whole blocks of `raw` -> (int16 audio, {tap: [per-block arrays]}, rds bits per block, events, stderr text)."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        nblk = raw.size // 307200
        audio, cap, bits, events, text = [], {t: [] for t in taps}, [], [], []
        for b in range(nblk):
            blk = self.block_id
            audio.append(self.block(raw[b * 307200:(b + 1) * 307200]))
            for t in taps:
                cap[t].append(self.tap(t))
            if self.mode != 1:
                bb, ev = self.rds()
                bits.append(bb)
                events.extend(ev)
                text.append(format_block(blk, self.rds_offset, ev))
        return np.concatenate(audio) if audio else np.zeros(0, np.int16), cap, bits, events, "".join(text)
