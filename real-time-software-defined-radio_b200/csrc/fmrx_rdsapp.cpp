// RDS data-link and application layer: block synchronisation, checkword verification with burst-error correction, group
// assembly, PI / PTY / TP / PS / RadioText decoding (IEC 62106).  Consumes the differentially decoded
// bits the GPU chain emits (the output of the reference's frame_thread up to src/fm_radio.cpp:616); the reference itself
// only prints syndrome matches (:625-718) and has no equivalent of this layer.  Host code, one record per station.
#include <cstdarg>
#include <cstring>
#include <new>
#include <vector>

#include "fmrx_internal.h"

namespace {

// parity-check matrix of the reference (src/fm_radio.cpp:477) as 10-bit rows, MSB = syndrome element 0: the syndrome of a
// 26-bit word (first received bit = row 0) is the XOR of the rows under its 1 bits
const uint16_t kH[26] = {0x200, 0x100, 0x080, 0x040, 0x020, 0x010, 0x008, 0x004, 0x002, 0x001, 0x2DC, 0x16E, 0x0B7,
                         0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};
// offset words A, B, C, C', D (added to the checkword) and the syndromes they leave on an error-free block
const uint16_t kOffsetWord[5] = {0x0FC, 0x198, 0x168, 0x350, 0x1B4};
enum { OFF_A = 0, OFF_B, OFF_C, OFF_CP, OFF_D };

uint16_t syndrome26(uint32_t w) {  // w: 26 bits, bit 25 = first received
    uint16_t s = 0;
    for (int i = 0; i < 26; ++i)
        if (w & (1u << (25 - i))) s ^= kH[i];
    return s;
}

struct Tables {
    uint16_t offset_syndrome[5];
    uint32_t burst[1024];  // error syndrome -> 26-bit error pattern (burst of <= 5 bits), 0 = none / ambiguous
    Tables() {
        for (int o = 0; o < 5; ++o) offset_syndrome[o] = syndrome26(kOffsetWord[o]);  // the offset sits in the check bits
        std::memset(burst, 0, sizeof(burst));
        bool clash[1024] = {};
        for (int len = 1; len <= 5; ++len)
            for (uint32_t pat = 1u << (len - 1); pat < (1u << len); ++pat) {
                if (!(pat & 1u)) continue;  // a burst of length `len` starts and ends with an error bit
                for (int sh = 0; sh + len <= 26; ++sh) {
                    const uint32_t e = pat << sh;
                    const uint16_t s = syndrome26(e);
                    if (burst[s] && burst[s] != e) clash[s] = true;
                    if (!burst[s]) burst[s] = e;
                }
            }
        for (int s = 0; s < 1024; ++s)
            if (clash[s]) burst[s] = 0;
    }
};
const Tables &tables() {
    static const Tables t;
    return t;
}

struct Station {
    uint32_t reg = 0;        // last 26 bits
    uint64_t nbits = 0;      // bits fed so far
    bool synced = false;
    int expect = 0;          // next block position 0..3 while synced
    uint64_t next_at = 0;    // value of nbits at which the next block completes
    int cand_pos = -1;       // unsynced: position (0..3) of the last syndrome hit and when it completed
    uint64_t cand_at = 0;
    uint16_t cand_info = 0;  // information word of that candidate block
    int bad_run = 0;
    // group under assembly
    uint16_t blk[4] = {0, 0, 0, 0};
    bool ok[4] = {false, false, false, false};
    bool is_cprime = false;
    int corrected_in_group = 0;
    uint64_t group_start = 0;
    // decoded programme data
    int pi = -1, pty = -1, tp = -1;
    char ps[8];
    uint8_t ps_mask = 0;
    char rt[64];
    uint64_t rt_mask = 0;
    int rt_ab = -1;
    uint32_t groups = 0, blocks_ok = 0, blocks_corrected = 0, blocks_bad = 0, sync_losses = 0;
    Station() { std::memset(ps, '_', sizeof(ps)); std::memset(rt, '_', sizeof(rt)); }
};

// which block position an offset index belongs to
inline int pos_of_offset(int o) { return o == OFF_A ? 0 : o == OFF_B ? 1 : o == OFF_D ? 3 : 2; }

void apply_group(Station &st, const fmrx_rds_group &g) {
    st.groups += 1;
    st.pi = g.blk[0];
    st.tp = (g.blk[1] >> 10) & 1;
    st.pty = (g.blk[1] >> 5) & 31;
    if (g.type == 0) {  // 0A / 0B: two PS characters in block D at segment address C1 C0
        const int seg = g.blk[1] & 3;
        st.ps[2 * seg] = (char)(g.blk[3] >> 8);
        st.ps[2 * seg + 1] = (char)(g.blk[3] & 0xFF);
        st.ps_mask |= (uint8_t)(3u << (2 * seg));
    } else if (g.type == 2) {  // 2A: four RadioText characters in blocks C, D; 2B: two in block D
        const int ab = (g.blk[1] >> 4) & 1, seg = g.blk[1] & 15;
        if (st.rt_ab != ab) {  // a toggled A/B flag announces a new message: clear the buffer
            std::memset(st.rt, '_', sizeof(st.rt));
            st.rt_mask = 0;
            st.rt_ab = ab;
        }
        if (!g.version_b) {
            const char c[4] = {(char)(g.blk[2] >> 8), (char)(g.blk[2] & 0xFF), (char)(g.blk[3] >> 8), (char)(g.blk[3] & 0xFF)};
            for (int i = 0; i < 4; ++i) st.rt[4 * seg + i] = c[i];
            st.rt_mask |= 15ull << (4 * seg);
        } else {
            st.rt[2 * seg] = (char)(g.blk[3] >> 8);
            st.rt[2 * seg + 1] = (char)(g.blk[3] & 0xFF);
            st.rt_mask |= 3ull << (2 * seg);
        }
    }
}

// one received bit; returns true when it completed a group (written to *out)
bool feed_bit(Station &st, int bit, fmrx_rds_group *out) {
    const Tables &T = tables();
    st.reg = ((st.reg << 1) | (uint32_t)(bit & 1)) & 0x3FFFFFFu;
    st.nbits += 1;
    if (st.nbits < 26) return false;
    if (!st.synced) {
        const uint16_t s = syndrome26(st.reg);
        for (int o = 0; o < 5; ++o) {
            if (s != T.offset_syndrome[o]) continue;
            const int pos = pos_of_offset(o);
            // two error-free blocks 26 bits apart in cyclic order A -> B -> C/C' -> D -> A acquire synchronisation
            if (st.cand_pos >= 0 && st.nbits - st.cand_at == 26 && pos == (st.cand_pos + 1) % 4) {
                st.synced = true;
                st.bad_run = 0;
                st.expect = (pos + 1) % 4;
                st.next_at = st.nbits + 26;
                for (bool &k : st.ok) k = false;
                st.corrected_in_group = 0;
                // the block just received already belongs to the group under assembly if the group started inside the
                // window (positions 0..pos): only a hit on A then B gives a complete group, so keep B when the candidate was A
                if (st.cand_pos == 0 && pos == 1) {
                    st.blk[0] = st.cand_info;  // see below: remembered with the candidate
                    st.ok[0] = true;
                    st.blk[1] = (uint16_t)(st.reg >> 10);
                    st.ok[1] = true;
                    st.group_start = st.cand_at - 26;
                    st.blocks_ok += 2;
                }
                st.cand_pos = -1;
                return false;
            }
            st.cand_pos = pos;
            st.cand_at = st.nbits;
            st.cand_info = (uint16_t)(st.reg >> 10);
            break;
        }
        return false;
    }
    if (st.nbits != st.next_at) return false;
    st.next_at += 26;
    const int pos = st.expect;
    st.expect = (pos + 1) % 4;
    const uint16_t s = syndrome26(st.reg);
    uint32_t word = st.reg;
    bool good = false, fixed = false, cprime = false;
    const int cands[2] = {pos == 0 ? OFF_A : pos == 1 ? OFF_B : pos == 2 ? OFF_C : OFF_D, pos == 2 ? OFF_CP : -1};
    for (int c = 0; c < 2 && !good; ++c) {
        if (cands[c] < 0) continue;
        if (s == T.offset_syndrome[cands[c]]) { good = true; cprime = cands[c] == OFF_CP; }
    }
    // correction is only trusted on a channel that is mostly clean: after three consecutive blocks that were not
    // received error-free, a syndrome that happens to match a burst pattern (40% of random words do) is not accepted
    for (int c = 0; c < 2 && !good && st.bad_run < 3; ++c) {
        if (cands[c] < 0) continue;
        const uint32_t e = T.burst[s ^ T.offset_syndrome[cands[c]]];
        if (e) { word ^= e; good = fixed = true; cprime = cands[c] == OFF_CP; }
    }
    if (pos == 0) {  // a new group starts: forget the previous one's blocks
        for (bool &k : st.ok) k = false;
        st.corrected_in_group = 0;
        st.group_start = st.nbits - 26;
    }
    if (good) {
        st.blk[pos] = (uint16_t)(word >> 10);
        st.ok[pos] = true;
        if (pos == 2) st.is_cprime = cprime;
        if (fixed) { st.blocks_corrected += 1; st.corrected_in_group += 1; }
        else st.blocks_ok += 1;
    } else {
        st.blocks_bad += 1;
    }
    if (good && !fixed) {
        st.bad_run = 0;
    } else {
        if (++st.bad_run >= 12) {  // three whole groups without a single error-free block: synchronisation is lost
            st.synced = false;
            st.sync_losses += 1;
            st.cand_pos = -1;
        }
    }
    if (pos == 3 && st.ok[0] && st.ok[1] && st.ok[2] && st.ok[3]) {
        fmrx_rds_group g{};
        for (int i = 0; i < 4; ++i) g.blk[i] = st.blk[i];
        g.type = (uint8_t)(st.blk[1] >> 12);
        g.version_b = (uint8_t)((st.blk[1] >> 11) & 1);
        g.corrected = (uint8_t)st.corrected_in_group;
        g.bit_index = (uint32_t)st.group_start;
        // a version-B group must carry C' and a version-A group C; a mismatch means a miscorrected block: drop the group
        if ((g.version_b != 0) != st.is_cprime) return false;
        apply_group(st, g);
        if (out) *out = g;
        return true;
    }
    return false;
}

}  // namespace

struct fmrx_rds_app {
    std::vector<Station> st;
};

extern "C" {

int fmrx_rds_app_create(int n_streams, fmrx_rds_app **out) {
    if (!out || n_streams <= 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_rds_app_create: bad argument");
    fmrx_rds_app *a = new (std::nothrow) fmrx_rds_app();
    if (!a) return fmrx::fail(FMRX_ERR_ALLOC, "out of host memory");
    a->st.resize((size_t)n_streams);
    *out = a;
    return FMRX_OK;
}

void fmrx_rds_app_destroy(fmrx_rds_app *a) { delete a; }

int fmrx_rds_app_reset(fmrx_rds_app *a) {
    if (!a) return fmrx::fail(FMRX_ERR_ARG, "null handle");
    for (auto &s : a->st) s = Station();
    return FMRX_OK;
}

int fmrx_rds_app_feed(fmrx_rds_app *a, const uint8_t *bits, const int32_t *n_bits, int n_blocks, fmrx_rds_group *groups, int cap, int32_t *n_groups) {
    if (!a || !bits || !n_bits || n_blocks <= 0 || cap < 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_rds_app_feed: bad argument");
    const size_t S = a->st.size();
    for (size_t s = 0; s < S; ++s) {
        Station &st = a->st[s];
        int ng = 0;
        for (int b = 0; b < n_blocks; ++b) {
            const int n = n_bits[s * n_blocks + b];
            if (n < 0 || n > FMRX_MAX_BITS) return fmrx::fail(FMRX_ERR_ARG, "n_bits[%zu][%d] = %d outside 0..%d", s, b, n, FMRX_MAX_BITS);
            const uint8_t *p = bits + (s * n_blocks + b) * FMRX_MAX_BITS;
            for (int i = 0; i < n; ++i) {
                fmrx_rds_group g;
                if (feed_bit(st, p[i], &g)) {
                    if (groups && ng < cap) groups[s * (size_t)cap + ng] = g;
                    ++ng;
                }
            }
        }
        if (n_groups) n_groups[s] = ng;
    }
    return FMRX_OK;
}

int fmrx_rds_app_station(const fmrx_rds_app *a, int stream, fmrx_rds_station *out) {
    if (!a || !out || stream < 0 || (size_t)stream >= a->st.size()) return fmrx::fail(FMRX_ERR_ARG, "fmrx_rds_app_station: bad argument");
    const Station &st = a->st[(size_t)stream];
    std::memset(out, 0, sizeof(*out));
    out->synced = st.synced ? 1 : 0;
    out->pi = st.pi; out->pty = st.pty; out->tp = st.tp;
    std::memcpy(out->ps, st.ps, 8);
    out->ps[8] = 0;
    size_t n = 64;
    for (size_t i = 0; i < 64; ++i)
        if (st.rt[i] == '\r') { n = i; break; }
    std::memcpy(out->rt, st.rt, n);
    out->rt[n] = 0;
    out->ps_complete = st.ps_mask == 0xFF;
    out->rt_ab_flag = (uint8_t)(st.rt_ab < 0 ? 0 : st.rt_ab);
    out->groups = st.groups; out->blocks_ok = st.blocks_ok; out->blocks_corrected = st.blocks_corrected; out->blocks_bad = st.blocks_bad;
    out->sync_losses = st.sync_losses;
    out->bits_fed = st.nbits;
    return FMRX_OK;
}

}  // extern "C"
