"""Mode 2 — 44.1 kHz audio through the reference's polyphase resampler with (U, D) = (147, 800) (BASELINE config 2).

The reference's main() only accepts mode 1 (src/fm_radio.cpp:736-764), so there is no reference output for this mode.
What pins it: (1) every function it is composed from is pinned against the reference (the 147/800 resampler included:
tests/test_oracle_golden.py::test_resamplers, golden `res_147_800`); (2) the composition is checked here — the oracle's
mode-2 chain must equal those pinned functions applied by hand to the mode-0 chain's signals, everything in front of the
audio resamplers must be bit-identical to mode 0, and the tones must come out where they were put at 44.1 kHz; (3) the
GPU chain and the executable must equal the oracle's mode-2 chain bit for bit.
"""
import subprocess

import numpy as np
import pytest

import fmrx
from fmrx import synth
from oracle import Chain, Port
from util import F, assert_bits

NA = 2822  # floor(15360 * 147 / 800)


@pytest.fixture(scope="module")
def port():
    return Port()


def test_oracle_mode2_is_the_pinned_functions_composed(port):
    nblk = 4
    raw = synth.synth_iq(nblk, 0, seed=3, f_l=1000.0, f_r=3000.0)
    c0, c2 = Chain(0, 1), Chain(2, 1)
    assert c2.n_audio == NA
    h = port.lpf(F(240000.0) * F(147.0), 16000.0, 151 * 147)  # src/fm_radio.cpp:200 with the mode-1 formula at U = 147
    assert h.size == 22197 and not np.isnan(h).any()  # odd length: no NaN tap (Q5)
    # states: sized taps - 1 as src/fm_radio.cpp:189-193 do, capped at N - 1 = 15359, the longest state the reference's
    # update rule zi[i] = x[N - Z - 1 + i] can fill from one block
    zm, zs = np.zeros(15359, F), np.zeros(15359, F)
    left, right = [], []
    for b in range(nblk):
        blk = raw[b * 307200:(b + 1) * 307200]
        c0.block(blk)
        a2 = c2.block(blk)
        for t in ("demod", "pilot", "nco", "stereo_bpf", "rds_rrc"):
            assert_bits(c2.tap(t), c0.tap(t), f"{t} block {b}: everything ahead of the audio resamplers is mode 0's")
        assert np.array_equal(c2.rds()[0], c0.rds()[0]) and c2.rds()[1] == c0.rds()[1]
        mono = port.resample(c0.tap("demod"), h, zm, 800, 147)
        mixed = (c0.tap("stereo_bpf") * c0.tap("nco")[:15360]).astype(F)
        st = port.resample(mixed, h, zs, 800, 147)
        assert_bits(c2.tap("mono"), mono, f"mono block {b}")
        assert_bits(c2.tap("stereo"), st, f"stereo block {b}")
        lr = c2.tap("audio_f").reshape(-1, 2)
        assert_bits(lr[:, 0], ((mono + st) / F(2)).astype(F), "L"); assert_bits(lr[:, 1], ((mono - st) / F(2)).astype(F), "R")
        q = (lr * F(16384.0) * F(147.0)).astype(np.int32).astype(np.int16)  # truncation toward zero, mult = U (src/fm_radio.cpp:290-298)
        assert np.array_equal(a2.reshape(-1, 2), q)
        if b >= 1:
            left.append(lr[:, 0] * 147.0); right.append(lr[:, 1] * 147.0)
    # the tones, at the new rate: L carries 1 kHz, R 3 kHz (the reference's half-weight stereo difference leaves a 3:1 mix)
    for sig, f0 in ((np.concatenate(left), 1000.0), (np.concatenate(right), 3000.0)):
        spec = np.abs(np.fft.rfft(sig * np.hanning(sig.size)))
        freqs = np.fft.rfftfreq(sig.size, 1 / 44100.0)
        band = (freqs > 200) & (freqs < 15000)
        peak = freqs[band][np.argmax(spec[band])]
        assert abs(peak - f0) < 12.0, (peak, f0)


@pytest.mark.gpu
@pytest.mark.parametrize("profile", [0, 1])
def test_gpu_mode2_equals_oracle(profile):
    nblk = 4
    raw = np.stack([synth.synth_iq(nblk, 0, seed=11 + s, f_l=700.0 + 300 * s, f_r=2500.0) for s in range(2)])
    with fmrx.Batch(2, mode=2, profile=profile, max_blocks=2) as rx:
        assert rx.n_audio == NA
        got = [rx.process(raw[:, k * 2 * 307200:(k + 1) * 2 * 307200].reshape(2, 2, 307200), want_float=True) for k in range(2)]  # two calls of two blocks: state carried
    for s in range(2):
        ch = Chain(2, profile)
        for b in range(nblk):
            ref = ch.block(raw[s, b * 307200:(b + 1) * 307200])
            res = got[b // 2]
            assert_bits(res["audio_f"][s, b % 2], ch.tap("audio_f"), f"float audio station {s} block {b}")
            assert np.array_equal(res["audio"][s, b % 2], ref), f"int16 station {s} block {b}"
            bits, _ = ch.rds()
            n = int(res["rds_n_bits"][s, b % 2])
            assert n == bits.size and np.array_equal(res["rds_bits"][s, b % 2, :n], bits), f"RDS bits station {s} block {b}"


@pytest.mark.gpu
def test_cli_audio_rate_44100():
    nblk = 3
    raw = synth.synth_iq(nblk, 0, seed=5)
    r = subprocess.run([fmrx.CLI_PATH, "--audio-rate", "44100", "--quiet"], input=raw.tobytes(), capture_output=True, timeout=600)
    assert r.returncode == 0, r.stderr.decode()
    audio = np.frombuffer(r.stdout, np.int16)
    assert audio.size == nblk * 2 * NA
    ch = Chain(2, 0)
    ref = np.concatenate([ch.block(raw[b * 307200:(b + 1) * 307200]) for b in range(nblk)])
    assert np.array_equal(audio, ref)
    bad = subprocess.run([fmrx.CLI_PATH, "1", "--audio-rate", "44100"], input=b"", capture_output=True, timeout=60)
    assert bad.returncode == 1  # 44.1 kHz is defined on the 2.4 Msps front end only
