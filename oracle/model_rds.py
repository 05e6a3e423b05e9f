"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the block loop of the reference's Python RDS model,
/root/reference/model/fmRDSblock.py:52-339 (numpy / scipy, float64, one station), each step citing the line it follows.

Parity status: PINNED.  tests/golden/make_model_rds_golden.py runs the unmodified script here (runpy, stub matplotlib) on a
seeded synthetic input and stores its stdout and per-block arrays in tests/golden/model_rds.npz; tests/test_model_rds.py
checks this port against that fixture (text identical, arrays bit-identical) and the GPU object fmrx.model_chain.ModelRds
against both.  bench.py times this port as the `cpu_baseline.python_models` leg on machines where the reference's model
directory does not exist (kind "port"); where it does, the script itself is timed (kind "reference").

The functions the script imports from its own directory (fmSupportLib.fmDemodArctan, fmPll.fmPll,
fmRRC.impulseResponseRootRaisedCosine) are restated below as well, so that the port has no import from the reference.
"""
import math

import numpy as np
from scipy import signal

# fmRDSblock.py:23-47
RF_FS, RF_FC, RF_TAPS, RF_DECIM = 2.4e6, 100e3, 151, 10
AUDIO_FS = 240000
BLOCK_SIZE = 307200  # :74

# parity-check matrix, fmRDSblock.py:50 (26 x 10), and the four syndromes it looks for, :286-313
H = np.array([[1, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0, 0, 0, 0, 0],
              [0, 0, 0, 0, 1, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 1, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1, 0, 0],
              [0, 0, 0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 1], [1, 0, 1, 1, 0, 1, 1, 1, 0, 0], [0, 1, 0, 1, 1, 0, 1, 1, 1, 0],
              [0, 0, 1, 0, 1, 1, 0, 1, 1, 1], [1, 0, 1, 0, 0, 0, 0, 1, 1, 1], [1, 1, 1, 0, 0, 1, 1, 1, 1, 1], [1, 1, 0, 0, 0, 1, 0, 0, 1, 1],
              [1, 1, 0, 1, 0, 1, 0, 1, 0, 1], [1, 1, 0, 1, 1, 1, 0, 1, 1, 0], [0, 1, 1, 0, 1, 1, 1, 0, 1, 1], [1, 0, 0, 0, 0, 0, 0, 0, 0, 1],
              [1, 1, 1, 1, 0, 1, 1, 1, 0, 0], [0, 1, 1, 1, 1, 0, 1, 1, 1, 0], [0, 0, 1, 1, 1, 1, 0, 1, 1, 1], [1, 0, 1, 0, 1, 0, 0, 1, 1, 1],
              [1, 1, 1, 0, 0, 0, 1, 1, 1, 1], [1, 1, 0, 0, 0, 1, 1, 0, 1, 1]], dtype=np.int64)
SYNDROMES = (("A", [1, 1, 1, 1, 0, 1, 1, 0, 0, 0]), ("B", [1, 1, 1, 1, 0, 1, 0, 1, 0, 0]), ("C", [1, 0, 0, 1, 0, 1, 1, 1, 0, 0]), ("D", [1, 0, 0, 1, 0, 1, 1, 0, 0, 0]))


def fm_demod_arctan(I, Q, prev_phase=0.0):
    """model/fmSupportLib.py:12-44: arctan2, numpy.unwrap of [prev, current], difference"""
    out = np.empty(len(I))
    for k in range(len(I)):
        cur = math.atan2(Q[k], I[k])
        prev_phase, cur = np.unwrap([prev_phase, cur])
        out[k] = cur - prev_phase
        prev_phase = cur  # the UNWRAPPED phase is what is carried (:33-40)
    return out, prev_phase


def fm_pll(pll_in, freq, Fs, state, nco_scale=1.0, phase_adjust=0.0, norm_bandwidth=0.01):
    """model/fmPll.py:4-56; state = [integrator, phaseEst, feedbackI, feedbackQ, ncoOut[0], trigOffset]"""
    Cp, Ci = 2.666, 3.555
    Kp, Ki = norm_bandwidth * Cp, norm_bandwidth * norm_bandwidth * Ci
    nco, nco_q = np.empty(len(pll_in) + 1), np.empty(len(pll_in) + 1)
    integrator, phase_est, fb_i, fb_q = state[0], state[1], state[2], state[3]
    nco[0], trig_offset = state[4], state[5]
    nco_q[0] = 0.0  # the model leaves ncoOutQ[0] uninitialised (np.empty); it is never read by the RDS script's decisions
    for k in range(len(pll_in)):
        err_i = pll_in[k] * (+fb_i)
        err_q = pll_in[k] * (-fb_q)
        err_d = math.atan2(err_q, err_i)
        integrator = integrator + Ki * err_d
        phase_est = phase_est + Kp * err_d + integrator
        trig_arg = 2 * math.pi * (freq / Fs) * (trig_offset + k + 1) + phase_est
        fb_i, fb_q = math.cos(trig_arg), math.sin(trig_arg)
        nco[k + 1] = math.cos(trig_arg * nco_scale + phase_adjust)
        nco_q[k + 1] = math.sin(trig_arg * nco_scale + phase_adjust)
    return nco, nco_q, [integrator, phase_est, fb_i, fb_q, nco[-1], trig_offset + len(pll_in)]


def rrc_taps(Fs, N_taps):
    """model/fmRRC.py: root-raised-cosine, T_symbol = 1/2375, beta = 0.90"""
    T, beta = 1 / 2375.0, 0.90
    h = np.empty(N_taps)
    for k in range(N_taps):
        t = float((k - N_taps / 2)) / Fs
        if t == 0.0:
            h[k] = 1.0 + beta * ((4 / math.pi) - 1)
        elif t == -T / (4 * beta) or t == T / (4 * beta):
            h[k] = (beta / np.sqrt(2)) * (((1 + 2 / math.pi) * (math.sin(math.pi / (4 * beta)))) + ((1 - 2 / math.pi) * (math.cos(math.pi / (4 * beta)))))
        else:
            h[k] = (math.sin(math.pi * t * (1 - beta) / T) + 4 * beta * (t / T) * math.cos(math.pi * t * (1 + beta) / T)) / (math.pi * t * (1 - (4 * beta * t / T) * (4 * beta * t / T)) / T)
    return h


class RdsDecisions:
    """fmRDSblock.py:205-337 for one station: clock pick (signed maximum in block 0, then re-derived from the position of the
    last symbol, :208-219), Manchester screening (:233-251), bit decisions with the stale-bit-free fresh array (:253-277),
    differential decoding (:281-292), the sliding syndrome check without a resync counter (:296-333)."""

    def __init__(self):
        self.block_count = 0
        self.int_offset = 0
        self.start_pos = 0
        self.lonely_bit = 0
        self.front_bit = 0
        self.prebit = 0
        self.printposition = 0
        self.prev_sync_bits = np.zeros(0)
        self.last_position = -1

    def block(self, rrc_rds):
        lines = []
        if self.block_count == 0:  # :207-213
            self.int_offset = int(np.where(rrc_rds[0:24] == np.max(rrc_rds[0:24]))[0][0])
            lines.append("Initial offset for clock recovery  %d" % self.int_offset)
        sampled_at = self.int_offset
        symbols_I = rrc_rds[self.int_offset::24]                                                                     # :216
        self.int_offset = int(24 - np.where(rrc_rds[len(rrc_rds) - 24::] == symbols_I[-1])[0][0])                    # :219
        if self.block_count == 0:  # :233-251
            c0 = c1 = 0
            for m in range(int(len(symbols_I) / 4)):
                if (symbols_I[2 * m] > 0 and symbols_I[2 * m + 1] > 0) or (symbols_I[2 * m] < 0 and symbols_I[2 * m + 1] < 0):
                    c0 += 1
                elif (symbols_I[2 * m + 1] > 0 and symbols_I[2 * m + 2] > 0) or (symbols_I[2 * m + 1] < 0 and symbols_I[2 * m + 2] < 0):
                    c1 += 1
            lines.append("Amount of doub when start 0  %d  Amount of doub when 1  %d" % (c0, c1))
            if c0 > c1:
                self.start_pos = 1
            elif c1 > c0:
                self.start_pos = 0
            lines.append("Start position  %d" % self.start_pos)
        sp = self.start_pos
        bit_stream = np.zeros(int(len(symbols_I) / 2) - sp)                                                          # :253
        if sp == 1 and self.block_count != 0:                                                                        # :257-261
            if self.lonely_bit > symbols_I[0]:
                self.front_bit = 1
            elif self.lonely_bit < symbols_I[0]:
                self.front_bit = 0
        for k in range(len(bit_stream)):                                                                             # :263-271
            if sp + 2 * k + 1 > len(symbols_I) - 1:
                break
            if symbols_I[2 * k + sp] > symbols_I[2 * k + 1 + sp]:
                bit_stream[k] = 1
            elif symbols_I[2 * k + sp] < symbols_I[2 * k + 1 + sp]:
                bit_stream[k] = 0
        if sp == 1:                                                                                                  # :273-277
            bit_stream = np.insert(bit_stream, 0, self.front_bit, axis=0)
            self.lonely_bit = symbols_I[-1]
        if self.block_count == 0:                                                                                    # :281-285
            self.prebit = bit_stream[0]
            offset = 1
        else:
            offset = 0
        diff_bits = np.zeros(len(bit_stream) - offset)
        for t in range(len(diff_bits)):                                                                              # :288-290
            diff_bits[t] = float(bool(self.prebit) != bool(bit_stream[t + offset]))
            self.prebit = bit_stream[t + offset]
        self.prebit = bit_stream[-1]                                                                                 # :292
        new_bits = diff_bits.copy()
        if self.block_count != 0:                                                                                    # :296-297
            diff_bits = np.insert(diff_bits, 0, self.prev_sync_bits, axis=0)
        position = 0
        events = []
        while True:                                                                                                  # :300-331
            block = diff_bits[position:position + 26].astype(np.int64)
            syn = (block @ H[:len(block)]) % 2 if len(block) == 26 else None
            for letter, pattern in SYNDROMES:
                if syn is not None and syn.tolist() == pattern:
                    if self.last_position == -1 or self.printposition - self.last_position == 26:
                        lines.append("Syndrome %s at position  %d" % (letter, self.printposition))
                        self.last_position = self.printposition
                        events.append((self.block_count, 0, "ABCD".index(letter), self.printposition))
                    else:
                        lines.append("False positive Syndrome %s at position  %d" % (letter, self.printposition))
                        events.append((self.block_count, 1, "ABCD".index(letter), self.printposition))
                    break
            position += 1
            if position + 26 > len(diff_bits) - 1:
                break
            self.printposition += 1
        self.prev_sync_bits = diff_bits[position - 1::]                                                              # :333
        self.block_count += 1
        return dict(symbols_I=symbols_I, sampled_at=sampled_at, bits=new_bits.astype(np.uint8), diff_bits=diff_bits, events=events, lines=lines)


class ModelRdsPort:
    """The whole loop body for one station: u8 block in, every array the script forms for that block out."""

    def __init__(self):
        nyq = AUDIO_FS / 2
        self.rf_coeff = signal.firwin(RF_TAPS, RF_FC / (RF_FS / 2), window=("hann"))                                              # :64
        self.extract_coeff = signal.firwin(RF_TAPS, [54000 / nyq, 60000 / nyq], window=("hann"), pass_zero="bandpass")            # :88
        self.square_coeff = signal.firwin(RF_TAPS, [113500 / nyq, 114500 / nyq], window=("hann"), pass_zero="bandpass")           # :91
        self.lpf_coeff = signal.firwin(RF_TAPS, 3000 / nyq, window=("hann"))                                                      # :99
        self.anti_coeff = signal.firwin(RF_TAPS, (57000 / 2) / ((240000 * 19) / 2), window=("hann"))                              # :105
        self.rrc_coeff = rrc_taps(57000, 151)                                                                                      # :111
        z = lambda: np.zeros(RF_TAPS - 1)  # noqa: E731
        self.st_i, self.st_q, self.st_extract, self.st_square = z(), z(), z(), z()
        self.st_lpf, self.st_lpf_q, self.st_anti, self.st_anti_q, self.st_rrc, self.st_rrc_q = z(), z(), z(), z(), z(), z()
        self.state_phase = 0.0
        self.phase_adj = math.pi / 3.3 - math.pi / 1.5                                                                            # :95
        self.state_pll = [0.0, 0.0, 1.0, 0.0, 1.0, 0.0]                                                                           # :96
        self.dec = RdsDecisions()

    def dsp(self, raw_u8):
        """:130-203 -> rrc_rds, rrc_rds_Q and the intermediates"""
        iq = (np.asarray(raw_u8, np.uint8) - 128.0) / 128.0                                                                        # :58-59
        i_filt, self.st_i = signal.lfilter(self.rf_coeff, 1.0, iq[0::2], zi=self.st_i)                                             # :130-135
        q_filt, self.st_q = signal.lfilter(self.rf_coeff, 1.0, iq[1::2], zi=self.st_q)
        i_ds, q_ds = i_filt[::RF_DECIM], q_filt[::RF_DECIM]
        fm_demod, self.state_phase = fm_demod_arctan(i_ds, q_ds, self.state_phase)                                                 # :142
        extract, self.st_extract = signal.lfilter(self.extract_coeff, 1.0, fm_demod, zi=self.st_extract)                           # :153
        squared = np.square(extract)                                                                                               # :158
        pre_pll, self.st_square = signal.lfilter(self.square_coeff, 1.0, squared, zi=self.st_square)                               # :161
        nco, nco_q, self.state_pll = fm_pll(pre_pll, 114000, 240000, self.state_pll, 0.5, self.phase_adj, 0.001)                   # :164
        mixed = np.multiply(extract, nco[0:len(extract):1]) * 2                                                                    # :170
        mixed_q = np.multiply(extract, nco_q[0:len(extract):1]) * 2                                                                # :172
        lpf, self.st_lpf = signal.lfilter(self.lpf_coeff, 1.0, mixed, zi=self.st_lpf)                                              # :177
        lpf_q, self.st_lpf_q = signal.lfilter(self.lpf_coeff, 1.0, mixed_q, zi=self.st_lpf_q)                                      # :179
        up, up_q = np.zeros(len(lpf) * 19), np.zeros(len(lpf) * 19)                                                                # :181-188
        up[::19], up_q[::19] = lpf, lpf_q
        anti, self.st_anti = signal.lfilter(self.anti_coeff, 1.0, up, zi=self.st_anti)                                             # :191
        anti_q, self.st_anti_q = signal.lfilter(self.anti_coeff, 1.0, up_q, zi=self.st_anti_q)                                     # :193
        res, res_q = anti[::80] * 19, anti_q[::80] * 19                                                                            # :195-196
        rrc, self.st_rrc = signal.lfilter(self.rrc_coeff, 1.0, res, zi=self.st_rrc)                                                # :199
        rrc_q, self.st_rrc_q = signal.lfilter(self.rrc_coeff, 1.0, res_q, zi=self.st_rrc_q)                                        # :201
        return dict(fm_demod=fm_demod, extract_rds=extract, pre_Pll_rds=pre_pll, post_Pll=nco, post_Pll_Q=nco_q, lpf_filt_rds=lpf, resample_rds=res,
                    rrc_rds=rrc, rrc_rds_Q=rrc_q)

    def block(self, raw_u8):
        out = self.dsp(raw_u8)
        out.update(self.dec.block(out["rrc_rds"]))
        out["symbols_Q"] = out["rrc_rds_Q"][out["sampled_at"]::24]                                                                 # :217
        return out


def run_script_blocks(raw_u8):
    """The script's loop bounds (:61, :127): at most 8 blocks of input are kept and the last block is never processed."""
    data = np.asarray(raw_u8, np.uint8)[:8 * BLOCK_SIZE]
    n = 0
    while (n + 1) * BLOCK_SIZE < len(data):
        n += 1
    return n


class ModelMonoPort:
    """model/fmMonoBlock.py:43-175 for one station (float IQ in [-1, 1], the script's 102400-value blocks): mono path, stereo
    carrier recovery, channel extraction, x2 mixer, stereo low-pass and the combiner as the script writes it -- its three
    names alias one array (:166-170), so what it stores in both channels is (audio - stereo) / 4.  Pinned by
    tests/golden/model_chain.npz (generated from the script's own statements and the reference's fmPll / fmDemodArctan)."""

    def __init__(self):
        nyq = 240e3 / 2
        self.rf_coeff = signal.firwin(151, 100e3 / (2.4e6 / 2), window=("hann"))                                               # :55
        self.audio_coeff = signal.firwin(151, 16e3 / nyq, window=("hann"))                                                    # :58
        self.bp = signal.firwin(151, [18.5e3 / nyq, 19.5e3 / nyq], window=("hann"), pass_zero="bandpass")                     # :115
        self.ext = signal.firwin(151, [22e3 / nyq, 54e3 / nyq], window=("hann"), pass_zero="bandpass")                        # :151
        self.st = [np.zeros(150) for _ in range(6)]
        self.phase = 0.0
        self.pll_state = [0.0, 0.0, 1.0, 0.0, 1.0, 0.0]

    def block(self, blk):
        st = self.st
        i_filt, st[0] = signal.lfilter(self.rf_coeff, 1.0, blk[0::2], zi=st[0])
        q_filt, st[1] = signal.lfilter(self.rf_coeff, 1.0, blk[1::2], zi=st[1])
        fm_demod, self.phase = fm_demod_arctan(i_filt[::10], q_filt[::10], self.phase)
        audio_filt, st[2] = signal.lfilter(self.audio_coeff, 1.0, fm_demod, zi=st[2])
        audio = audio_filt[::5].copy()
        pilot, st[3] = signal.lfilter(self.bp, 1.0, fm_demod, zi=st[3])
        nco, _, self.pll_state = fm_pll(pilot, 19e3, 240e3, self.pll_state, 2)
        ext, st[4] = signal.lfilter(self.ext, 1.0, fm_demod, zi=st[4])
        mixed = np.multiply(nco[0:len(ext):1], ext) * 2
        stereo_filt, st[5] = signal.lfilter(self.audio_coeff, 1.0, mixed, zi=st[5])
        stereo = stereo_filt[::5]
        combined = ((audio + stereo) / 2 - stereo) / 2
        return dict(audio=audio, pilot=pilot, nco=nco, stereo=stereo, combined=combined, fm_demod=fm_demod)
