"""Synthetic FM broadcast multiplex -> 8-bit interleaved IQ (the input format of `fm_radio`, src/fm_radio.cpp:66).

This is the generator BASELINE.json's north_star asks for: known tone, pilot and RDS content, 75 kHz deviation,
quantised to u8 at the mode's RF rate (2.4 Msps mode 0 / 2.5 Msps mode 1, src/fm_radio.cpp:36-37).  It follows
SURVEY App. E with one change: RDS chips are shaped with a root-raised-cosine pulse (so that the receiver's RRC matched
filter, src/filter.cpp:63-93, completes a raised cosine and the symbol decisions have wide margins).

The same code runs on numpy (parity tests, golden fixtures) and on torch tensors (bench: thousands of stations are
synthesised directly in HBM).  Only elementwise ops, cumsum and a table gather are used.
"""
from __future__ import annotations

import numpy as np

BLOCK_BYTES = 307200  # src/fm_radio.cpp:23
BLOCK_IQ = BLOCK_BYTES // 2
CHIP_RATE = 2375.0  # 2 chips per RDS bit at 1187.5 bit/s
RDS_OFFSETS = (0x0FC, 0x198, 0x168, 0x1B4)  # offset words A, B, C, D
RDS_POLY = 0x5B9  # g(x) = x^10 + x^8 + x^7 + x^5 + x^4 + x^3 + 1
_PULSE_SPAN = 4  # RRC pulse support, chips either side
_PULSE_RES = 2048  # table points per chip
# Chip timing offset.  The reference picks its symbol sampling phase ONCE, from the filter start-up transient of block 0
# (src/fm_radio.cpp:503-517), which always lands on sample 23 of 24; the transmitter therefore has to
# put its chip centres there.  10/24 of a chip gives BER 0 and the widest decision margins through the reference.
RDS_T0 = (10.0 / 24.0) / CHIP_RATE


def rf_rate(mode: int) -> float:
    return 2.5e6 if mode == 1 else 2.4e6


def rds_checkword(info16: int, offset: int) -> int:
    """10-bit RDS checkword: remainder of info*x^10 by g(x), plus the offset word."""
    reg = info16 << 10
    for bit in range(25, 9, -1):
        if reg & (1 << bit):
            reg ^= RDS_POLY << (bit - 10)
    return (reg & 0x3FF) ^ offset


def rds_bits(n_bits: int, seed: int) -> np.ndarray:
    """Valid RDS groups (4 blocks x (16 info + 10 check)), MSB first, random payload from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    n_blocks = (n_bits + 25) // 26 + 1
    out = np.empty(n_blocks * 26, dtype=np.uint8)
    infos = rng.integers(0, 1 << 16, size=n_blocks)
    for b in range(n_blocks):
        word = (int(infos[b]) << 10) | rds_checkword(int(infos[b]), RDS_OFFSETS[b % 4])
        for i in range(26):
            out[b * 26 + i] = (word >> (25 - i)) & 1
    return out[:n_bits]


def rds_block_bits(info16: int, offset: int) -> list:
    word = ((info16 & 0xFFFF) << 10) | rds_checkword(info16 & 0xFFFF, offset)
    return [(word >> (25 - i)) & 1 for i in range(26)]


RDS_OFFSET_CPRIME = 0x350


def rds_group_bits(pi: int, ps: str = "FMRX-GPU", rt: str = "", pty: int = 10, tp: int = 0, n_bits: int = 0, version_b: bool = False,
                   rt_ab: int = 0) -> np.ndarray:
    """Bit stream (MSB first, before differential encoding) of a programme that cycles through its four 0A (or 0B) groups
    carrying the 8-character PS name and, if `rt` is given, its 2A (or 2B) RadioText groups, repeated to `n_bits`
    (one pass if 0).  IEC 62106 group layout: block B = type(4) version(1) TP(1) PTY(5) + 5 group-specific bits."""
    ps = (ps + " " * 8)[:8]
    groups = []
    for seg in range(4):
        b = (0 << 12) | (int(version_b) << 11) | (tp << 10) | (pty << 5) | (1 << 2) | seg  # TA 0, M/S 0, DI bit 1
        c = pi if version_b else 0xE0CD  # 0A block C: alternative frequencies (filler code pair); 0B: PI again
        groups.append((pi, b, c, (ord(ps[2 * seg]) << 8) | ord(ps[2 * seg + 1])))
    if rt:
        per = 2 if version_b else 4
        text = rt if len(rt) >= (32 if version_b else 64) else rt + "\r"
        text = text + " " * (-len(text) % per)
        for seg in range(len(text) // per):
            b = (2 << 12) | (int(version_b) << 11) | (tp << 10) | (pty << 5) | (rt_ab << 4) | seg
            ch = [ord(x) for x in text[per * seg:per * seg + per]]
            if version_b:
                groups.append((pi, b, pi, (ch[0] << 8) | ch[1]))
            else:
                groups.append((pi, b, (ch[0] << 8) | ch[1], (ch[2] << 8) | ch[3]))
    one = []
    for a, b, c, d in groups:
        one += rds_block_bits(a, RDS_OFFSETS[0]) + rds_block_bits(b, RDS_OFFSETS[1]) + rds_block_bits(c, RDS_OFFSET_CPRIME if version_b else RDS_OFFSETS[2]) + rds_block_bits(d, RDS_OFFSETS[3])
    one = np.array(one, dtype=np.uint8)
    if n_bits <= 0:
        return one
    return np.tile(one, n_bits // len(one) + 1)[:n_bits]


def rds_chips(bits: np.ndarray) -> np.ndarray:
    """Differential encoding d[i] = d[i-1]^b[i], then biphase: 1 -> (+1,-1), 0 -> (-1,+1)."""
    d = np.bitwise_xor.accumulate(bits.astype(np.uint8))
    s = 2.0 * d.astype(np.float64) - 1.0
    chips = np.empty(2 * len(d), dtype=np.float64)
    chips[0::2] = s
    chips[1::2] = -s
    return chips


def _rrc_pulse_table(beta: float = 0.9) -> np.ndarray:
    """Unit-energy-free RRC pulse g(u), u = t/Tc in [-SPAN, SPAN], sampled at _PULSE_RES points per chip, peak 1."""
    u = np.arange(-_PULSE_SPAN * _PULSE_RES, _PULSE_SPAN * _PULSE_RES + 1, dtype=np.float64) / _PULSE_RES
    g = np.empty_like(u)
    sing = np.isclose(np.abs(u), 1.0 / (4.0 * beta))
    zero = np.isclose(u, 0.0)
    reg = ~(sing | zero)
    ur = u[reg]
    g[reg] = (np.sin(np.pi * ur * (1 - beta)) + 4 * beta * ur * np.cos(np.pi * ur * (1 + beta))) / (np.pi * ur * (1 - (4 * beta * ur) ** 2))
    g[zero] = 1.0 + beta * (4 / np.pi - 1)
    g[sing] = (beta / np.sqrt(2.0)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
    return g / g.max()


_PULSE = None


def _pulse():
    global _PULSE
    if _PULSE is None:
        _PULSE = _rrc_pulse_table()
    return _PULSE


def rds_baseband(chips: np.ndarray, t: np.ndarray) -> np.ndarray:
    """s(t) = sum_c chips[c] * g(t*CHIP_RATE - c); chips outside the array count as 0."""
    g = _pulse()
    u = t * CHIP_RATE
    c0 = np.floor(u).astype(np.int64)
    out = np.zeros_like(t)
    for dc in range(-_PULSE_SPAN + 1, _PULSE_SPAN + 1):
        c = c0 + dc
        idx = np.rint((u - c + _PULSE_SPAN) * _PULSE_RES).astype(np.int64)
        ok = (c >= 0) & (c < len(chips)) & (idx >= 0) & (idx < len(g))
        out += np.where(ok, chips[np.clip(c, 0, len(chips) - 1)] * g[np.clip(idx, 0, len(g) - 1)], 0.0)
    return out


def station_params(s: int) -> dict:
    """Per-station content for the batch configuration."""
    return dict(seed=s + 1, f_l=400.0 + 37.0 * (s % 64), f_r=900.0 + 53.0 * (s % 64))


def synth_iq(n_blocks: int, mode: int = 0, seed: int = 1, f_l: float = 1000.0, f_r: float = 3000.0, rds: bool = True,
             rds_level: float = 0.05, pilot_level: float = 0.08, audio_level: float = 0.45, rds_payload=None, cnr_db=None,
             noise_seed: int = 0) -> np.ndarray:
    """Returns n_blocks*307200 bytes of interleaved u8 I,Q.  `rds_payload`: a callable n_bits -> bit array (e.g. a
    programme from rds_group_bits) replacing the random groups of `seed`.  `cnr_db`: carrier-to-noise ratio of white
    Gaussian noise added to I and Q ahead of the 8-bit quantiser (carrier power 1, noise power 2 sigma^2 over the full
    RF rate; None = noiseless), drawn from default_rng(noise_seed)."""
    fs = rf_rate(mode)
    n = n_blocks * BLOCK_IQ
    t = np.arange(n, dtype=np.float64) / fs
    left = 0.5 * np.sin(2 * np.pi * f_l * t)
    right = 0.5 * np.sin(2 * np.pi * f_r * t)
    th = 2 * np.pi * 19000.0 * t
    m = audio_level * (left + right) + audio_level * (left - right) * np.cos(2 * th) + pilot_level * np.cos(th)
    if rds:
        n_chips = int(np.ceil(t[-1] * CHIP_RATE)) + 2 * _PULSE_SPAN + 2
        bits = rds_payload((n_chips + 1) // 2) if rds_payload is not None else rds_bits((n_chips + 1) // 2, seed)
        m = m + rds_level * rds_baseband(rds_chips(bits), t - RDS_T0) * np.cos(3 * th)
    phi = 2 * np.pi * 75e3 * np.cumsum(m) / fs
    out = np.empty(2 * n, dtype=np.uint8)
    ci, cq = np.cos(phi), np.sin(phi)
    if cnr_db is not None:
        sigma = np.sqrt(0.5 * 10.0 ** (-float(cnr_db) / 10.0))
        rng = np.random.default_rng(noise_seed)
        ci = ci + sigma * rng.standard_normal(n)
        cq = cq + sigma * rng.standard_normal(n)
    out[0::2] = np.clip(np.rint(127.0 * ci + 128.0), 0, 255).astype(np.uint8)
    out[1::2] = np.clip(np.rint(127.0 * cq + 128.0), 0, 255).astype(np.uint8)
    return out


def synth_station(s: int, n_blocks: int, mode: int = 0) -> np.ndarray:
    return synth_iq(n_blocks, mode=mode, **station_params(s))


def synth_batch_torch(stations, n_blocks: int, mode: int, device, chunk: int = 64, cnr_db=None, noise_seed: int = 0):
    """Same multiplex for many stations at once, synthesised on `device` with torch (float64).  Returns a uint8
    tensor [len(stations), n_blocks*307200].  Used by bench.py to fill HBM with thousands of distinct stations in
    seconds; parity spot checks copy individual rows back and run the oracle on those exact bytes.  `cnr_db`: white Gaussian
    noise ahead of the quantiser as in synth_iq (torch's generator, seeded with noise_seed: not numpy's noise, the same statistics)."""
    import torch

    fs = rf_rate(mode)
    n = n_blocks * BLOCK_IQ
    stations = list(stations)
    out = torch.empty((len(stations), 2 * n), dtype=torch.uint8, device=device)
    t = torch.arange(n, dtype=torch.float64, device=device) / fs
    th = 2 * np.pi * 19000.0 * t
    c2, c1, c3 = torch.cos(2 * th), torch.cos(th), torch.cos(3 * th)
    g = torch.as_tensor(_pulse(), device=device)
    u = (t - RDS_T0) * CHIP_RATE
    c0 = torch.floor(u).to(torch.int64)
    n_chips = int(np.ceil(float(t[-1]) * CHIP_RATE)) + 2 * _PULSE_SPAN + 2
    for lo in range(0, len(stations), chunk):
        ids = stations[lo:lo + chunk]
        par = [station_params(s) for s in ids]
        fl = torch.tensor([p["f_l"] for p in par], dtype=torch.float64, device=device)[:, None]
        fr = torch.tensor([p["f_r"] for p in par], dtype=torch.float64, device=device)[:, None]
        left = 0.5 * torch.sin(2 * np.pi * fl * t[None, :])
        right = 0.5 * torch.sin(2 * np.pi * fr * t[None, :])
        m = 0.45 * (left + right) + 0.45 * (left - right) * c2[None, :] + 0.08 * c1[None, :]
        chips = torch.as_tensor(np.stack([rds_chips(rds_bits((n_chips + 1) // 2, p["seed"])) for p in par]), device=device)
        bb = torch.zeros_like(m)
        for dc in range(-_PULSE_SPAN + 1, _PULSE_SPAN + 1):
            c = c0 + dc
            idx = torch.round((u - c + _PULSE_SPAN) * _PULSE_RES).to(torch.int64)
            ok = (c >= 0) & (c < chips.shape[1]) & (idx >= 0) & (idx < g.numel())
            w = torch.where(ok, g[idx.clamp(0, g.numel() - 1)], torch.zeros((), dtype=torch.float64, device=device))
            bb += chips[:, c.clamp(0, chips.shape[1] - 1)] * w[None, :]
        m = m + 0.05 * bb * c3[None, :]
        phi = 2 * np.pi * 75e3 * torch.cumsum(m, dim=1) / fs
        row = out[lo:lo + len(ids)]
        ci, cq = torch.cos(phi), torch.sin(phi)
        if cnr_db is not None:
            gen = torch.Generator(device=device)
            gen.manual_seed(int(noise_seed) * 1000003 + lo)
            sigma = float(np.sqrt(0.5 * 10.0 ** (-float(cnr_db) / 10.0)))
            ci = ci + sigma * torch.randn(ci.shape, dtype=torch.float64, device=device, generator=gen)
            cq = cq + sigma * torch.randn(cq.shape, dtype=torch.float64, device=device, generator=gen)
        row[:, 0::2] = torch.clamp(torch.round(127.0 * ci + 128.0), 0, 255).to(torch.uint8)
        row[:, 1::2] = torch.clamp(torch.round(127.0 * cq + 128.0), 0, 255).to(torch.uint8)
    return out
