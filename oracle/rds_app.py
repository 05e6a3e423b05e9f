"""TEST INFRASTRUCTURE ONLY.  Independent Python restatement of the RDS data-link / application layer that
csrc/fmrx_rdsapp.cpp implements (IEC 62106: offset words, checkword by polynomial division with g(x) = 0x5B9,
burst-error correction of bursts <= 5 bits, group assembly, PI / PTY / TP / PS / RadioText).  The reference has no such
layer (it stops at syndrome print-outs, src/fm_radio.cpp:625-718), so this oracle is pinned by known-answer vectors:
groups encoded from known PI / PS / RadioText by fmrx.synth.rds_group_bits and decoded back.

Synchronisation policy (shared with the C++ by specification, include/fmrx.h): acquire on two error-free blocks 26 bits
apart in cyclic order A -> B -> C/C' -> D -> A; lose it after 12 consecutive blocks that were not received error-free;
attempt burst correction only while fewer than 3 consecutive blocks failed the clean check."""
POLY = 0x5B9
OFFSETS = {"A": 0x0FC, "B": 0x198, "C": 0x168, "C'": 0x350, "D": 0x1B4}
POS = {"A": 0, "B": 1, "C": 2, "C'": 2, "D": 3}


def remainder(word26: int) -> int:
    """remainder of word26(x) modulo g(x): zero for info*x^10 + checkword(info)"""
    reg = word26
    for bit in range(25, 9, -1):
        if reg >> bit & 1:
            reg ^= POLY << (bit - 10)
    return reg & 0x3FF


def _burst_table():
    tab, clash = {}, set()
    for length in range(1, 6):
        for pat in range(1 << (length - 1), 1 << length):
            if not pat & 1:
                continue
            for sh in range(0, 27 - length):
                e = pat << sh
                s = remainder(e)
                if s in tab and tab[s] != e:
                    clash.add(s)
                tab.setdefault(s, e)
    for s in clash:
        del tab[s]
    return tab


BURST = _burst_table()


class Station:
    def __init__(self):
        self.reg, self.n, self.synced = 0, 0, False
        self.expect, self.next_at, self.cand, self.bad_run = 0, 0, None, 0
        self.blk, self.ok, self.cprime, self.ncorr, self.start = [0] * 4, [False] * 4, False, 0, 0
        self.pi = self.pty = self.tp = -1
        self.ps, self.rt, self.rt_ab = ["_"] * 8, ["_"] * 64, -1
        self.groups, self.blocks_ok, self.blocks_corrected, self.blocks_bad, self.sync_losses = [], 0, 0, 0, 0

    def _apply(self, g):
        b = g["blk"]
        self.pi, self.tp, self.pty = b[0], b[1] >> 10 & 1, b[1] >> 5 & 31
        if g["type"] == 0:
            seg = b[1] & 3
            self.ps[2 * seg], self.ps[2 * seg + 1] = chr(b[3] >> 8), chr(b[3] & 255)
        elif g["type"] == 2:
            ab, seg = b[1] >> 4 & 1, b[1] & 15
            if self.rt_ab != ab:
                self.rt, self.rt_ab = ["_"] * 64, ab
            if not g["version_b"]:
                self.rt[4 * seg:4 * seg + 4] = [chr(b[2] >> 8), chr(b[2] & 255), chr(b[3] >> 8), chr(b[3] & 255)]
            else:
                self.rt[2 * seg:2 * seg + 2] = [chr(b[3] >> 8), chr(b[3] & 255)]

    def feed(self, bits):
        out = []
        for bit in bits:
            self.reg = ((self.reg << 1) | (int(bit) & 1)) & 0x3FFFFFF
            self.n += 1
            if self.n < 26:
                continue
            rem = remainder(self.reg)
            if not self.synced:
                for name, off in OFFSETS.items():
                    if rem != off:  # remainder of a valid block = its offset word
                        continue
                    pos = POS[name]
                    if self.cand and self.n - self.cand[1] == 26 and pos == (self.cand[0] + 1) % 4:
                        self.synced, self.bad_run, self.expect, self.next_at = True, 0, (pos + 1) % 4, self.n + 26
                        self.ok, self.ncorr = [False] * 4, 0
                        if self.cand[0] == 0 and pos == 1:
                            self.blk[0], self.blk[1], self.ok[0], self.ok[1] = self.cand[2], self.reg >> 10, True, True
                            self.start = self.cand[1] - 26
                            self.blocks_ok += 2
                        self.cand = None
                    else:
                        self.cand = (pos, self.n, self.reg >> 10)
                    break
                continue
            if self.n != self.next_at:
                continue
            self.next_at += 26
            pos, self.expect = self.expect, (self.expect + 1) % 4
            names = [("A",), ("B",), ("C", "C'"), ("D",)][pos]
            word, good, fixed, cprime = self.reg, False, False, False
            for nm in names:
                if rem == OFFSETS[nm]:
                    good, cprime = True, nm == "C'"
                    break
            if not good and self.bad_run < 3:
                for nm in names:
                    e = BURST.get(rem ^ OFFSETS[nm])
                    if e:
                        word, good, fixed, cprime = word ^ e, True, True, nm == "C'"
                        break
            if pos == 0:
                self.ok, self.ncorr, self.start = [False] * 4, 0, self.n - 26
            if good:
                self.blk[pos], self.ok[pos] = word >> 10, True
                if pos == 2:
                    self.cprime = cprime
                if fixed:
                    self.blocks_corrected += 1
                    self.ncorr += 1
                else:
                    self.blocks_ok += 1
            else:
                self.blocks_bad += 1
            if good and not fixed:
                self.bad_run = 0
            else:
                self.bad_run += 1
                if self.bad_run >= 12:
                    self.synced, self.cand = False, None
                    self.sync_losses += 1
            if pos == 3 and all(self.ok):
                g = dict(blk=list(self.blk), type=self.blk[1] >> 12, version_b=self.blk[1] >> 11 & 1, corrected=self.ncorr, bit_index=self.start)
                if bool(g["version_b"]) != self.cprime:
                    continue
                self._apply(g)
                self.groups.append(g)
                out.append(g)
        return out

    @property
    def ps_text(self):
        return "".join(self.ps)

    @property
    def rt_text(self):
        return "".join(self.rt).split("\r")[0]
