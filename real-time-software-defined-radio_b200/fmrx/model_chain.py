"""The block loops of the reference's Python models (model/fmMonoBlock.py:43-175, model/fmRDSblock.py:52-339) as objects over the GPU's
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


class ModelRds:
    """The block loop of model/fmRDSblock.py (:127-337) for `S` stations in lock step: the signal path -- front end, 54-60 kHz
    extraction, squaring + 113.5-114.5 kHz band-pass, fmPll at 114 kHz with its in-phase AND quadrature outputs, x2 mixers, 3 kHz
    low-pass, zero-stuff x19 / anti-image / [::80] x19, RRC, on both branches -- runs on the GPU's model-compatible operators
    (float64); the decisions (1187.5 bit/s per station) are host integer logic that follows the script statement by statement:
    the sampling phase is the SIGNED maximum of the first 24 RRC samples and is re-derived at the end of every block from
    where the last symbol sat (:208-219), bits are decided into a fresh zero array (a tie is a 0, :253-271), the syndrome
    check has no false-positive counter and never re-synchronises (:300-331).  This is numerically and logically a different
    receiver from the C++ one; `text` reproduces the script's stdout for the block."""

    BLOCK_BYTES = 307200
    # parity-check matrix rows as 10-bit words (fmRDSblock.py:50) and the four syndromes the script compares with (:286-313)
    H_ROWS = (0x200, 0x100, 0x080, 0x040, 0x020, 0x010, 0x008, 0x004, 0x002, 0x001, 0x2DC, 0x16E, 0x0B7, 0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201,
              0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B)
    SYNDROMES = {0x3D8: "A", 0x3D4: "B", 0x25C: "C", 0x258: "D"}

    def __init__(self, n_streams=1):
        from . import model_firwin

        S, nyq, T = n_streams, 240000 / 2, 151
        self.S = S
        self.rf_coeff = model_firwin(T, 100e3 / (2.4e6 / 2))                                   # :64
        self.extract_coeff = model_firwin(T, [54000 / nyq, 60000 / nyq], pass_zero=False)     # :88
        self.square_coeff = model_firwin(T, [113500 / nyq, 114500 / nyq], pass_zero=False)    # :91
        self.lpf_coeff = model_firwin(T, 3000 / nyq)                                          # :99
        self.anti_coeff = model_firwin(T, (57000 / 2) / ((240000 * 19) / 2))                  # :105
        self.rrc_coeff = self._rrc(57000.0, T)                                                # :111
        z = lambda: np.zeros((S, T - 1))  # noqa: E731
        self.st = {k: z() for k in ("i", "q", "extract", "square", "lpf", "lpf_q", "anti", "anti_q", "rrc", "rrc_q")}
        self.state_phase = np.zeros(S)
        self.phase_adj = np.pi / 3.3 - np.pi / 1.5                                            # :95
        self.state_pll = np.tile(np.array([0.0, 0.0, 1.0, 0.0, 1.0, 0.0]), (S, 1))            # :96
        self.block_count = 0
        self.dec = [dict(int_offset=0, start_pos=0, lonely=0.0, front_bit=0, prebit=0, printposition=0, prev=np.zeros(0, np.uint8), last_position=-1)
                    for _ in range(S)]

    @staticmethod
    def _rrc(Fs, n):
        """model/fmRRC.py: T_symbol = 1/2375, beta = 0.90, the 1/T scale factor ignored; float64"""
        T, beta = 1 / 2375.0, 0.90
        t = (np.arange(n) - n / 2) / Fs
        with np.errstate(divide="ignore", invalid="ignore"):
            h = (np.sin(np.pi * t * (1 - beta) / T) + 4 * beta * (t / T) * np.cos(np.pi * t * (1 + beta) / T)) / (np.pi * t * (1 - (4 * beta * t / T) * (4 * beta * t / T)) / T)
        h[t == 0.0] = 1.0 + beta * ((4 / np.pi) - 1)
        sing = (t == -T / (4 * beta)) | (t == T / (4 * beta))
        h[sing] = (beta / np.sqrt(2)) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
        return h

    def signal_path(self, raw):
        """u8 [S][307200] -> the script's per-block arrays up to rrc_rds / rrc_rds_Q ([S][3648])"""
        from . import model_demod, model_lfilter, model_pll

        S, st = self.S, self.st
        iq = (np.asarray(raw, np.uint8).reshape(S, self.BLOCK_BYTES) - 128.0) / 128.0                         # :58-59
        i_ds = model_lfilter(iq[:, 0::2], self.rf_coeff, st["i"], decim=10).reshape(S, -1)                    # :130-139
        q_ds = model_lfilter(iq[:, 1::2], self.rf_coeff, st["q"], decim=10).reshape(S, -1)
        fm_demod = model_demod(i_ds, q_ds, self.state_phase).reshape(S, -1)                                   # :142
        extract = model_lfilter(fm_demod, self.extract_coeff, st["extract"]).reshape(S, -1)                   # :153
        pre_pll = model_lfilter(np.square(extract), self.square_coeff, st["square"]).reshape(S, -1)           # :158-161
        nco, nco_q = model_pll(pre_pll, 114000, 240000, self.state_pll, 0.5, self.phase_adj, 0.001)           # :164
        nco, nco_q = nco.reshape(S, -1), nco_q.reshape(S, -1)
        n = extract.shape[1]
        out = dict(fm_demod=fm_demod, extract_rds=extract, pre_Pll_rds=pre_pll, post_Pll=nco, post_Pll_Q=nco_q)
        for tag, osc in (("", nco), ("_Q", nco_q)):
            k = tag.lower()
            mixed = extract * osc[:, :n] * 2                                                                  # :170-172
            lpf = model_lfilter(mixed, self.lpf_coeff, st["lpf" + k]).reshape(S, -1)                          # :177-179
            res = model_lfilter(lpf, self.anti_coeff, st["anti" + k], decim=80, up=19).reshape(S, -1) * 19    # :181-196: only the retained outputs
            out["lpf_filt_rds" + tag], out["resample_rds" + tag] = lpf, res
            out["rrc_rds" + tag] = model_lfilter(res, self.rrc_coeff, st["rrc" + k]).reshape(S, -1)           # :199-201
        return out

    def _decide(self, d, rrc, rrc_q):
        first = self.block_count == 0
        lines = []
        if first:
            d["int_offset"] = int(np.argmax(rrc[:24]))                       # the first position of the signed maximum, :208
            lines.append("Initial offset for clock recovery  %d" % d["int_offset"])
        at = d["int_offset"]
        sym, sym_q = rrc[at::24], rrc_q[at::24]                              # :216-217
        d["int_offset"] = 24 - int(np.flatnonzero(rrc[-24:] == sym[-1])[0])  # :219
        if first:                                                            # :233-251
            a, b, c = sym[0:2 * (len(sym) // 4):2], sym[1:2 * (len(sym) // 4) + 1:2], sym[2:2 * (len(sym) // 4) + 2:2]
            same01 = ((a > 0) & (b > 0)) | ((a < 0) & (b < 0))
            same12 = ((b > 0) & (c > 0)) | ((b < 0) & (c < 0))
            c0, c1 = int(np.count_nonzero(same01)), int(np.count_nonzero(~same01 & same12))
            lines.append("Amount of doub when start 0  %d  Amount of doub when 1  %d" % (c0, c1))
            if c0 != c1:
                d["start_pos"] = 1 if c0 > c1 else 0
            lines.append("Start position  %d" % d["start_pos"])
        sp = d["start_pos"]
        nb = len(sym) // 2 - sp                                              # :253
        k = np.arange(nb)
        k = k[sp + 2 * k + 1 <= len(sym) - 1]                                # :265-266
        bits = np.zeros(nb, np.uint8)
        bits[k] = sym[2 * k + sp] > sym[2 * k + 1 + sp]                      # a tie leaves the fresh zero, :268-271
        if sp == 1:
            if not first and d["lonely"] != sym[0]:                          # :257-261
                d["front_bit"] = int(d["lonely"] > sym[0])
            bits = np.concatenate([[d["front_bit"]], bits]).astype(np.uint8)  # :276
            d["lonely"] = float(sym[-1])
        if first:                                                            # :281-285
            d["prebit"], bits_in = int(bits[0]), bits[1:]
        else:
            bits_in = bits
        diff = (np.concatenate([[d["prebit"]], bits_in[:-1]]) ^ bits_in).astype(np.uint8) if bits_in.size else np.zeros(0, np.uint8)  # :288-290
        d["prebit"] = int(bits[-1])                                          # :292
        stream = diff if first else np.concatenate([d["prev"], diff])       # :296-297
        events, position = [], 0
        while True:                                                          # :300-331
            syn = 0
            for j in np.flatnonzero(stream[position:position + 26]):
                syn ^= self.H_ROWS[j]
            letter = self.SYNDROMES.get(syn)
            if letter is not None:
                good = d["last_position"] == -1 or d["printposition"] - d["last_position"] == 26
                lines.append(("Syndrome %s at position  %d" if good else "False positive Syndrome %s at position  %d") % (letter, d["printposition"]))
                events.append((self.block_count, 0 if good else 1, "ABCD".index(letter), d["printposition"]))
                if good:
                    d["last_position"] = d["printposition"]
            position += 1
            if position + 26 > len(stream) - 1:
                break
            d["printposition"] += 1
        d["prev"] = stream[position - 1:]                                    # :333
        return dict(symbols_I=sym, symbols_Q=sym_q, bits=diff, events=events, lines=lines)

    def block(self, raw):
        """raw: u8 [S][307200] (or [307200] for one station).  Returns the signal-path arrays plus, per station, `symbols_I`,
        `symbols_Q`, `bits` (the differentially decoded bits this block added), `events` ((block, 0 good / 1 false positive,
        letter, position)) and `text`: the lines the script prints for this block."""
        out = self.signal_path(raw)
        per = [self._decide(self.dec[s], out["rrc_rds"][s], out["rrc_rds_Q"][s]) for s in range(self.S)]
        for key in ("symbols_I", "symbols_Q", "bits", "events"):
            out[key] = [p[key] for p in per]
        out["text"] = ["\nProcessing block %d\n" % self.block_count + "".join(line + "\n" for line in p["lines"]) for p in per]
        self.block_count += 1
        return out
