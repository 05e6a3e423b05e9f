/* TEST INFRASTRUCTURE ONLY — see fmrx_oracle.h for the rules and the parity status (PINNED against the reference).
 *
 * Plain-C restatement of the reference algorithm, written to be bit-comparable with the reference as built by
 * `g++ -O3` on x86-64 (SSE2 scalar float, no FMA, no excess precision): every operand type, promotion and rounding
 * point below follows the reference expression it cites.  Compile with -ffp-contract=off (oracle/Makefile).
 * All file:line citations are relative to /root/reference/.
 */
#include "fmrx_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846 /* src/dy4.h:13 */

/* ------------------------------------------------------------------------------------------------------------ */
/* filter design                                                                                                 */
/* ------------------------------------------------------------------------------------------------------------ */

/* src/filter.cpp:19-38.  norm_cutoff is fp32; the centre test uses (N-1)/2 but the sinc argument uses the integer
 * N/2 (:29 vs :33), so an even N yields 0/0 = NaN at i = N/2.  Window = sin^2(i*pi/N). */
void orc_design_lpf(float Fs, float Fc, unsigned short ntaps, float *h) {
    float nc = (float)((double)Fc / ((double)Fs / 2.0));
    int N = ntaps;
    for (int i = 0; i < N; i++) {
        float v;
        if (i == (N - 1) / 2) {
            v = nc;
        } else {
            double d = (double)(i - N / 2);
            double arg = ORC_PI * (double)nc * d;
            float sinc = (float)(sin(arg) / arg);
            v = nc * sinc;
        }
        double w = sin(((double)i * ORC_PI) / (double)N);
        h[i] = (float)((double)v * (w * w));
    }
}

/* src/filter.cpp:41-60.  All of norm_pass, n_half, norm_center are fp32; n_half comes from an integer division. */
void orc_design_bpf(float Fb, float Fe, float Fs, int ntaps, float *h) {
    float half_fs = Fs / 2.0f;
    float np = (Fe - Fb) / half_fs;
    float n_half = (float)((ntaps - 1) / 2);
    float ncen = ((Fe + Fb) / 2.0f) / half_fs;
    for (int i = 0; i < ntaps; i++) {
        float v;
        if ((float)i == n_half) {
            v = np;
        } else {
            double off = (double)((float)i - n_half);
            double arg = ORC_PI * (double)(np / 2.0f) * off;
            v = (float)((double)np * sin(arg) / arg);
        }
        v = (float)((double)v * cos((double)i * ORC_PI * (double)ncen));
        double w = sin(((double)i * ORC_PI) / (double)ntaps);
        h[i] = (float)((double)v * (w * w));
    }
}

/* src/filter.cpp:63-93.  t is fp32, everything else double; beta = 0.90f, T = (float)(1/2375). */
void orc_design_rrc(float Fs, int ntaps, float *h) {
    float T = (float)(1.0 / 2375.0);
    float beta = 0.90f;
    double Td = (double)T, bd = (double)beta;
    for (int k = 0; k < ntaps; k++) {
        float t = (float)((double)k - (double)ntaps / 2.0) / Fs;
        double td = (double)t;
        double v;
        if (td == 0.0) {
            v = 1.0 + bd * ((4 / ORC_PI) - 1);
        } else if (td == -Td / (4.0 * bd) || td == Td / (4.0 * bd)) {
            v = (bd / sqrt(2.0)) * (((1.0 + 2.0 / ORC_PI) * (sin(ORC_PI / (4.0 * bd)))) + ((1.0 - 2.0 / ORC_PI) * (cos(ORC_PI / (4.0 * bd)))));
        } else {
            double a = 4.0 * bd * td / Td;
            double num = sin(ORC_PI * td * (1.0 - bd) / Td) + 4.0 * bd * (double)(t / T) * cos(ORC_PI * td * (1.0 + bd) / Td);
            double den = ORC_PI * td * (1.0 - a * a) / Td;
            v = num / den;
        }
        h[k] = (float)v;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* per-sample primitives                                                                                         */
/* ------------------------------------------------------------------------------------------------------------ */

/* src/iofunc.cpp:67 — (u8 - 128)/128.0 in double, exact in fp32 */
void orc_unpack(const uint8_t *raw, size_t n, float *dst) {
    for (size_t k = 0; k < n; k++) dst[k] = (float)(((int)raw[k] - 128) / 128.0);
}

/* src/filter.cpp:126-154 (and :157-185).  History X(-j) = zi[nzi-j]; state saved one sample late (Q1). */
int orc_fir_decim(float *y, const float *x, int n, const float *h, int ntaps, float *zi, int nzi, int decim) {
    int ny = n / decim;
    for (int o = 0; o < ny; o++) {
        float acc = 0.0f;
        int count = 0;
        for (int k = 0; k < ntaps; k++) {
            int p = decim * o - k;
            if (p >= 0 && p < n) {
                acc = acc + x[p] * h[k];
            } else {
                acc = acc + zi[nzi - 1 - count] * h[k];
                count++;
            }
        }
        y[o] = acc;
    }
    for (int i = 0; i < nzi; i++) zi[i] = x[n - nzi - 1 + i];
    return ny;
}

/* src/filter.cpp:187-219 */
int orc_fir_decim_iq(float *yi, float *yq, const float *xi, const float *xq, int n, const float *h, int ntaps,
                     float *zii, float *ziq, int decim) {
    int nzi = ntaps - 1;
    int ny = n / decim;
    for (int o = 0; o < ny; o++) {
        float ai = 0.0f, aq = 0.0f;
        int count = 0;
        for (int k = 0; k < ntaps; k++) {
            int p = decim * o - k;
            if (p >= 0 && p < n) {
                ai = ai + xi[p] * h[k];
                aq = aq + xq[p] * h[k];
            } else {
                ai = ai + zii[nzi - 1 - count] * h[k];
                aq = aq + ziq[nzi - 1 - count] * h[k];
                count++;
            }
        }
        yi[o] = ai;
        yq[o] = aq;
    }
    for (int i = 0; i < nzi; i++) {
        zii[i] = xi[n - nzi - 1 + i];
        ziq[i] = xq[n - nzi - 1 + i];
    }
    return ny;
}

/* src/filter.cpp:222-339.  Taps k = k0 + c*up with k0 = (decim*o) mod up; in-range taps read x[(decim*o-k)/up], the
 * others read zi[(nzi-1-c)/up] where c counts EVERY visited tap (Q6).  The RDS variant scales by `up` (:333). */
int orc_resample(float *y, int ny_limit, const float *x, int n, const float *h, int ntaps, float *zi, int nzi,
                 int decim, int up, int gain_up) {
    long ny_full = ((long)n * up) / decim;
    int ny = (ny_limit > 0 && ny_limit < ny_full) ? ny_limit : (int)ny_full;
    long span = (long)n * up;
    for (int o = 0; o < ny; o++) {
        float acc = 0.0f;
        int count = 0;
        long base = (long)decim * o;
        for (long k = base % up; k < ntaps; k += up) {
            long p = base - k;
            if (p >= 0 && p < span) {
                acc = acc + x[p / up] * h[k];
            } else {
                acc = acc + zi[(nzi - 1 - count) / up] * h[k];
            }
            count++;
        }
        if (gain_up) acc = acc * (float)up;
        y[o] = acc;
    }
    for (int i = 0; i < nzi; i++) zi[i] = x[n - nzi - 1 + i];
    return (int)ny_full;
}

/* src/filter.cpp:373-401 with the geometry of its only call site (src/fm_radio.cpp:404): the reference sizes the
 * loop from the (n+1)-long NCO vector, so its state update reads products n-150 .. n-1 — a correct (not one-late)
 * history — but stores them WITHOUT the x2 the in-range taps get (Q8).  Output element n (which reads sig[n] one past
 * the end in the reference) is never consumed and is not produced here. */
int orc_fir_mixer(float *y, const float *nco, const float *sig, int n, const float *h, int ntaps, float *zi) {
    int nzi = ntaps - 1;
    for (int o = 0; o < n; o++) {
        float acc = 0.0f;
        int count = 0;
        for (int k = 0; k < ntaps; k++) {
            int p = o - k;
            if (p >= 0) {
                acc = acc + nco[p] * sig[p] * h[k] * 2.0f;
            } else {
                acc = acc + zi[nzi - 1 - count] * h[k];
                count++;
            }
        }
        y[o] = acc;
    }
    for (int i = 0; i < nzi; i++) zi[i] = nco[n - nzi + i] * sig[n - nzi + i];
    return n;
}

/* src/rf_module.cpp:13-34 — derivative-form discriminator; previous sample zeroed on entry (Q3); numerator fp32,
 * denominator pow(float,2) -> double. */
void orc_demod(const float *I, const float *Q, int n, float *dst) {
    float pi_ = 0.0f, pq_ = 0.0f;
    for (int k = 0; k < n; k++) {
        double den = (double)I[k] * (double)I[k] + (double)Q[k] * (double)Q[k];
        if (den == 0) {
            dst[k] = 0.0f;
        } else {
            float num = I[k] * (Q[k] - pq_) - Q[k] * (I[k] - pi_);
            dst[k] = (float)((double)num / den);
        }
        pi_ = I[k];
        pq_ = Q[k];
    }
}

/* one PLL step: src/helper.cpp:34-44 (identical at :149-159).  State/locals fp32; atan2/cos/sin are the double libm
 * versions; trigArg is a double expression rounded to fp32 (Q13). */
typedef struct { float integ, phase, fbi, fbq, off; float Ki, Kp, fratio, scale, adj; } pll_ctx;

static inline float pll_step(pll_ctx *c, float in, int k) {
    float eI = in * (+c->fbi);
    float eQ = in * (-c->fbq);
    float eD = (float)atan2((double)eQ, (double)eI);
    c->integ = c->integ + c->Ki * eD;
    c->phase = c->phase + (c->Kp * eD + c->integ);
    float cnt = (c->off + (float)k) + 1.0f;
    float trig = (float)((2 * ORC_PI) * (double)c->fratio * (double)cnt + (double)c->phase);
    c->fbi = (float)cos((double)trig);
    c->fbq = (float)sin((double)trig);
    return (float)cos((double)(trig * c->scale + c->adj));
}

static void pll_load(pll_ctx *c, const float *st, float freq, float Fs, float scale, float adj, float bw) {
    float Cp = 2.666f, Ci = 3.555f; /* src/helper.cpp:15-16 */
    c->Ki = (bw * bw) * Ci;
    c->Kp = bw * Cp;
    c->integ = st[0]; c->phase = st[1]; c->fbi = st[2]; c->fbq = st[3]; c->off = st[4];
    c->fratio = freq / Fs;
    c->scale = scale;
    c->adj = adj;
}

/* src/helper.cpp:13-57.  Output sample k is the NCO value computed at step k-1 (nco[0] = carried ncoLast). */
void orc_pll(float *nco, const float *x, int n, float freq, float Fs, float scale, float phase_adj, float bw, float *st) {
    pll_ctx c;
    pll_load(&c, st, freq, Fs, scale, phase_adj, bw);
    float last = st[5];
    for (int k = 0; k < n; k++) {
        nco[k] = last;
        last = pll_step(&c, x[k], k);
    }
    st[0] = c.integ; st[1] = c.phase; st[2] = c.fbi; st[3] = c.fbq; st[4] = c.off + (float)n; st[5] = last;
}

/* src/helper.cpp:108-173.  BPF on x^2 with the product formed in double (`pow(x,2)*h[k]`, :139) and accumulated into
 * an fp32 y; history holds (float)x^2, one-late (:162-164); then one PLL step per output.  nco is n+1 long. */
void orc_pll_combine(float *y, float *nco, const float *x, int n, const float *h, int ntaps, float *zi, float freq,
                     float Fs, float scale, float phase_adj, float bw, float *st) {
    pll_ctx c;
    pll_load(&c, st, freq, Fs, scale, phase_adj, bw);
    int nzi = ntaps - 1;
    nco[0] = st[5];
    for (int o = 0; o < n; o++) {
        float acc = 0.0f;
        int count = 0;
        for (int k = 0; k < ntaps; k++) {
            int p = o - k;
            if (p >= 0) {
                acc = (float)((double)acc + ((double)x[p] * (double)x[p]) * (double)h[k]);
            } else {
                acc = acc + zi[nzi - 1 - count] * h[k];
                count++;
            }
        }
        y[o] = acc;
        nco[o + 1] = pll_step(&c, acc, o);
    }
    for (int i = 0; i < nzi; i++) zi[i] = (float)((double)x[n - nzi - 1 + i] * (double)x[n - nzi - 1 + i]);
    st[0] = c.integ; st[1] = c.phase; st[2] = c.fbi; st[3] = c.fbq; st[4] = c.off + (float)n; st[5] = nco[n];
}

/* ------------------------------------------------------------------------------------------------------------ */
/* RDS clock/data recovery + frame sync: src/fm_radio.cpp:444-729                                                */
/* ------------------------------------------------------------------------------------------------------------ */

#define RDS_SPS 24      /* samples per chip at 57 kHz (:519-526) */
#define RDS_MAXSYM 4096

struct orc_rds_decoder {
    int block_id;
    unsigned initial_offset;
    unsigned start_pos;
    float lonely;      /* last unused chip when start_pos == 1 (:591) */
    int front_bit;     /* :458 */
    int prebit;        /* :460 */
    int bits[RDS_MAXSYM];   /* `bit_stream`: persistent, so ties keep stale values (Q12) */
    int nbits;
    int carry[27];     /* `prev_sync_bits` (:715-718) */
    unsigned printpos; /* `printposition` */
    int last_pos;      /* `last_position` */
    int bad;           /* `bad_sync_count` */
};

orc_rds_decoder *orc_rds_decoder_create(void) {
    orc_rds_decoder *d = (orc_rds_decoder *)calloc(1, sizeof(*d));
    d->last_pos = -1;
    return d;
}
void orc_rds_decoder_destroy(orc_rds_decoder *d) { free(d); }
int orc_rds_initial_offset(const orc_rds_decoder *d) { return (int)d->initial_offset; }
int orc_rds_start_pos(const orc_rds_decoder *d) { return (int)d->start_pos; }

/* Parity-check matrix rows as 10-bit words, MSB = syndrome element 0 (src/fm_radio.cpp:477) and the four offset-word
 * syndromes (:479-482).  Rows 0-9 are the identity. */
static const uint16_t RDS_H[26] = {0x200, 0x100, 0x080, 0x040, 0x020, 0x010, 0x008, 0x004, 0x002, 0x001, 0x2DC, 0x16E, 0x0B7,
                                   0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};
static const uint16_t RDS_SYN[4] = {0x3D8, 0x3D4, 0x25C, 0x258};

static int same_sign(float a, float b) { return (a > 0 && b > 0) || (a < 0 && b < 0); }

int orc_rds_decode_block(orc_rds_decoder *d, const float *rrc, int n, uint8_t *bits_out, orc_rds_event *ev, int ev_cap, int *n_ev) {
    float sym[RDS_MAXSYM];
    int nsym = n / RDS_SPS;
    int nev = 0;
    /* :503-517 — sampling phase picked once, from |rrc[0..23]| of block 0 (Q11) */
    if (d->block_id == 0) {
        float best = fabsf(rrc[0]);
        for (unsigned i = 1; i < RDS_SPS; i++) {
            if (fabsf(rrc[i]) > best) { best = fabsf(rrc[i]); d->initial_offset = i; }
        }
    }
    for (int k = 0; k < nsym; k++) sym[k] = rrc[RDS_SPS * k + d->initial_offset];
    /* :542-558 — Manchester alignment screening, block 0 only; loop index normalised to start at 0 (Q10) */
    if (d->block_id == 0) {
        int c0 = 0, c1 = 0;
        for (unsigned j = 0; j < (unsigned)nsym / 4; j++) {
            if (same_sign(sym[2 * j], sym[2 * j + 1])) c0++;
            else if (same_sign(sym[2 * j + 1], sym[2 * j + 2])) c1++;
        }
        if (c0 > c1) d->start_pos = 1;
        else if (c1 > c0) d->start_pos = 0;
    }
    /* :560 — resize keeps old contents, new elements are 0 */
    int want = nsym / 2 - (int)d->start_pos;
    for (int i = d->nbits; i < want; i++) d->bits[i] = 0;
    d->nbits = want;
    /* :565-572 */
    if (d->start_pos == 1 && d->block_id != 0) {
        if (d->lonely > sym[0]) d->front_bit = 1;
        else if (sym[0] > d->lonely) d->front_bit = 0;
    }
    /* :574-585 */
    for (int k = 0; k < d->nbits; k++) {
        unsigned a = 2u * (unsigned)k + d->start_pos;
        if (a + 1 > (unsigned)nsym - 1) break;
        if (sym[a] > sym[a + 1]) d->bits[k] = 1;
        else if (sym[a] < sym[a + 1]) d->bits[k] = 0;
    }
    /* :587-592 */
    if (d->start_pos == 1) {
        memmove(d->bits + 1, d->bits, sizeof(int) * (size_t)d->nbits);
        d->bits[0] = d->front_bit;
        d->nbits++;
        d->lonely = sym[nsym - 1];
    }
    /* :596-616 — differential decode; block 0 consumes its first bit as the reference */
    int off = 0;
    if (d->block_id == 0) { d->prebit = d->bits[0]; off = 1; }
    int diff[27 + RDS_MAXSYM];
    int ncarry = d->block_id != 0 ? 27 : 0;
    for (int g = 0; g < ncarry; g++) diff[g] = d->carry[g];
    int nd = d->nbits - off;
    for (int t = 0; t < nd; t++) {
        diff[ncarry + t] = d->prebit ^ d->bits[t + off];
        d->prebit = d->bits[t + off];
        if (bits_out) bits_out[t] = (uint8_t)diff[ncarry + t];
    }
    d->prebit = d->bits[d->nbits - 1];
    int total = ncarry + nd;
    /* :630-713 — sliding 26-bit syndrome */
    unsigned pos = 0;
    for (;;) {
        uint16_t s = 0;
        for (int j = 0; j < 26; j++) if (diff[pos + j]) s ^= RDS_H[j];
        for (int L = 0; L < 4; L++) {
            if (s != RDS_SYN[L]) continue;
            int good = (d->last_pos == -1) || (d->printpos - (unsigned)d->last_pos == 26u);
            if (nev < ev_cap) { ev[nev].block = d->block_id; ev[nev].kind = good ? ORC_EV_GOOD : ORC_EV_FALSE; ev[nev].letter = L; ev[nev].position = d->printpos; }
            nev++;
            if (good) { d->last_pos = (int)d->printpos; d->bad = 0; }
            else d->bad++;
            break;
        }
        if (d->bad > 10) {
            if (nev < ev_cap) { ev[nev].block = d->block_id; ev[nev].kind = ORC_EV_RESYNC; ev[nev].letter = -1; ev[nev].position = d->printpos; }
            nev++;
            d->bad = 0;
            d->last_pos = -1;
        }
        pos += 1;
        if (pos + 26 > (unsigned)total - 1) break;
        d->printpos += 1;
    }
    for (int g = 0; g < 27; g++) d->carry[g] = diff[pos - 1 + g];
    d->block_id++;
    if (n_ev) *n_ev = nev;
    return nd;
}

int orc_rds_format_block(int block_id, int initial_offset, const orc_rds_event *ev, int n_ev, char *buf, int cap) {
    int w = 0;
#define EMIT(...) do { int r_ = snprintf(buf + (w < cap ? w : cap), (size_t)(w < cap ? cap - w : 0), __VA_ARGS__); if (r_ > 0) w += r_; } while (0)
    if (block_id == 0) EMIT("initial offset for clock recovery = %d\n", initial_offset);
    EMIT(" \n****************Prcoessing Block: %d****************\n", block_id);
    for (int i = 0; i < n_ev; i++) {
        if (ev[i].kind == ORC_EV_RESYNC) EMIT("~~~~~Re-Sync~~~~~\n");
        else EMIT("%sSyndrome %c at position %u\n", ev[i].kind == ORC_EV_FALSE ? "False positive " : "", "ABCD"[ev[i].letter], ev[i].position);
    }
#undef EMIT
    return w;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* the chain                                                                                                     */
/* ------------------------------------------------------------------------------------------------------------ */

#define NIQ (ORC_BLOCK_BYTES / 2) /* 153600 complex samples per block (src/fm_radio.cpp:23) */
#define NIF (NIQ / 10)            /* 15360 */
#define NT 151

struct orc_chain {
    int mode, profile, block_id, paths;
    /* rf_thread (src/fm_radio.cpp:31-147) */
    float h_rf[NT], zi_i[NT - 1], zi_q[NT - 1];
    float *iq, *I, *Q, *If, *Qf, *demod;
    /* mono_stero_thread (:150-318) */
    int audio_taps, nzi_a, n_audio, audio_up, audio_decim, mult;
    float *h_mono, *h_stereo, h_pilot[NT], h_sbpf[NT];
    float *zi_mono, *zi_pilot, *zi_sbpf, *zi_stereo;
    float pll_st[6];
    float *mono, *pilot, *nco, *sbpf, *mixed, *stereo, *audio_f;
    /* rds_thread (:321-441) */
    float h_rbpf[NT], h_sq[NT], h_lpf3k[NT], h_rrc[NT], *h_anti;
    float zi_rbpf[NT - 1], zi_sq[NT - 1], zi_lpf[NT - 1], zi_rrc[NT - 1], *zi_anti;
    float rds_pll_st[6], rds_phase;
    float *rbpf, *rsq, *rnco, *rlpf, *rres, *rrrc;
    int n_rds;
    /* quality profile (not in the reference: SURVEY 8f row 4; restated from include/fmrx.h FMRX_QUALITY_*) */
    int quality, mono_delay;
    float mono_tail[16];
    double de_b, de_a1, de_state[4]; /* x[-1], y[-1] of L, then of R */
    /* frame_thread (:444-729) */
    orc_rds_decoder *dec;
    uint8_t bits[128];
    int nbits;
    orc_rds_event ev[256];
    int nev;
};

static float *fz(size_t n) { return (float *)calloc(n, sizeof(float)); }

orc_chain *orc_chain_create(int mode, int profile) {
    orc_chain *c = (orc_chain *)calloc(1, sizeof(*c));
    c->mode = mode; c->profile = profile; c->paths = 3;
    float rf_Fs = mode == 1 ? 2500000.0f : 2400000.0f; /* :36-37 */
    orc_design_lpf(rf_Fs, 100000.0f, NT, c->h_rf);     /* :40-42,75 */
    c->iq = fz(ORC_BLOCK_BYTES); c->I = fz(NIQ); c->Q = fz(NIQ); c->If = fz(NIF); c->Qf = fz(NIF); c->demod = fz(NIF);
    /* :153-180 */
    float audio_Fs = 240000.0f;
    c->audio_taps = NT; c->audio_up = 1; c->audio_decim = 5; c->mult = 1;
    if (mode == 1) { audio_Fs = 6000000.0f; c->audio_decim = 125; c->audio_up = 24; c->audio_taps = NT * 24; c->mult = 24; }
    /* mode 2 is NOT in the reference (its main() accepts only "1", src/fm_radio.cpp:736-764): the 44.1 kHz output BASELINE.json's
     * config 2 names, composed from the reference's own mode-1 thread body (:174-180, :228, :245) with (U, D) = (147, 800) on
     * the 240 kHz IF of mode 0 -- 151 * 147 = 22197 taps (odd: no NaN tap, Q5), floor(15360 * 147 / 800) = 2822 samples per
     * block.  Every function it calls is pinned against the reference; the composition has no reference output to pin. */
    if (mode == 2) { audio_Fs = 240000.0f * 147.0f; c->audio_decim = 800; c->audio_up = 147; c->audio_taps = NT * 147; c->mult = 147; }
    c->nzi_a = c->audio_taps - 1; /* all four states sized from audio_taps (:189-193) */
    if (c->nzi_a > NIF - 1) c->nzi_a = NIF - 1; /* mode 2 only: the update rule zi[i] = x[N - Z - 1 + i] needs Z <= N - 1 */
    c->n_audio = (int)(((long)NIF * c->audio_up) / c->audio_decim);
    c->h_mono = fz((size_t)c->audio_taps); c->h_stereo = fz((size_t)c->audio_taps);
    orc_design_lpf(audio_Fs, 16000.0f, (unsigned short)c->audio_taps, c->h_mono);   /* :200 */
    /* :201-202 design the two band-pass filters at audio_Fs, which in mode 1 is the UPSAMPLED rate (6 MHz) although they run
     * on the 250 kHz IF -- replicated for mode 1; mode 2 designs them at the rate they run at, like mode 0 */
    float bpf_Fs = mode == 2 ? 240000.0f : audio_Fs;
    orc_design_bpf(18.5e3f, 19.5e3f, bpf_Fs, NT, c->h_pilot);                        /* :201 */
    orc_design_bpf(22e3f, 54e3f, bpf_Fs, NT, c->h_sbpf);                             /* :202 */
    orc_design_lpf(audio_Fs, 16000.0f, (unsigned short)c->audio_taps, c->h_stereo); /* :203 */
    c->zi_mono = fz((size_t)c->nzi_a); c->zi_pilot = fz((size_t)c->nzi_a); c->zi_sbpf = fz((size_t)c->nzi_a); c->zi_stereo = fz((size_t)c->nzi_a);
    static const float pll0[6] = {0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 1.0f}; /* :165-171, :343-349 */
    memcpy(c->pll_st, pll0, sizeof(pll0)); memcpy(c->rds_pll_st, pll0, sizeof(pll0));
    c->mono = fz(NIF); c->pilot = fz(NIF); c->nco = fz(NIF + 1); c->sbpf = fz(NIF); c->mixed = fz(NIF); c->stereo = fz(NIF);
    c->audio_f = fz(2 * NIF);
    /* :331-370 */
    float Fs = 240000.0f;
    orc_design_bpf(54000.0f, 60000.0f, Fs, NT, c->h_rbpf);
    orc_design_bpf(113500.0f, 114500.0f, Fs, NT, c->h_sq);
    orc_design_lpf(Fs, 3000.0f, NT, c->h_lpf3k);
    c->h_anti = fz(NT * 19);
    orc_design_lpf(Fs * 19.0f, (float)(57000 / 2), NT * 19, c->h_anti);
    orc_design_rrc(57000.0f, NT, c->h_rrc);
    c->zi_anti = fz(NT * 19 - 1);
    float phase_adj = (float)(ORC_PI / 3.3 - ORC_PI / 1.5);      /* :342 */
    c->rds_phase = (float)((double)phase_adj - ORC_PI / 1.4);    /* :400 */
    c->rbpf = fz(NIF); c->rsq = fz(NIF); c->rnco = fz(NIF + 1); c->rlpf = fz(NIF + 1);
    c->n_rds = (int)(((long)(NIF + 1) * 19) / 80);
    c->rres = fz((size_t)c->n_rds); c->rrrc = fz((size_t)c->n_rds);
    c->dec = orc_rds_decoder_create();
    return c;
}

void orc_chain_destroy(orc_chain *c) {
    if (!c) return;
    float *p[] = {c->iq, c->I, c->Q, c->If, c->Qf, c->demod, c->h_mono, c->h_stereo, c->zi_mono, c->zi_pilot, c->zi_sbpf, c->zi_stereo,
                  c->mono, c->pilot, c->nco, c->sbpf, c->mixed, c->stereo, c->audio_f, c->h_anti, c->zi_anti, c->rbpf, c->rsq, c->rnco,
                  c->rlpf, c->rres, c->rrrc};
    for (size_t i = 0; i < sizeof(p) / sizeof(p[0]); i++) free(p[i]);
    orc_rds_decoder_destroy(c->dec);
    free(c);
}

/* response of a real FIR at f: the double sums the product's host code forms (csrc/fmrx_design.cpp fmrx_fir_response) */
static void fir_response(const float *h, int n, double Fs, double f, double *mag, double *phase) {
    double re = 0.0, im = 0.0, w = 2.0 * ORC_PI * f / Fs;
    for (int k = 0; k < n; k++) { re += (double)h[k] * cos(w * k); im -= (double)h[k] * sin(w * k); }
    if (mag) *mag = hypot(re, im);
    if (phase) *phase = atan2(im, re);
}
static void unity(float *h, int n, float Fb, float Fe, float Fs) {
    double mag;
    fir_response(h, n, (double)Fs, (double)((Fb + Fe) / 2.0f), &mag, NULL);
    for (int k = 0; k < n; k++) h[k] = (float)((double)h[k] / mag);
}

/* flags as FMRX_QUALITY_*: 1 / 2 de-emphasis 75 / 50 us, 4 unity-gain band-pass filters + x2 stereo mixer + delayed mono, 8 computed
 * RDS phase adjust.  Call right after orc_chain_create. */
void orc_chain_set_quality(orc_chain *c, int flags) {
    c->quality = flags;
    float bpf_Fs = c->mode == 1 ? 6000000.0f : 240000.0f;
    if (flags & 4) {
        unity(c->h_pilot, NT, 18.5e3f, 19.5e3f, bpf_Fs);
        unity(c->h_sbpf, NT, 22e3f, 54e3f, bpf_Fs);
        unity(c->h_rbpf, NT, 54000.0f, 60000.0f, 240000.0f);
        unity(c->h_sq, NT, 113500.0f, 114500.0f, 240000.0f);
        for (int k = 0; k < c->audio_taps; k++) c->h_stereo[k] *= 2.0f;
        c->mono_delay = (int)((75L * c->audio_up + c->audio_decim / 2) / c->audio_decim);
    }
    if (flags & 8) {
        double ph;
        fir_response(c->h_sq, NT, 240000.0, 114000.0, NULL, &ph);
        c->rds_phase = (float)(-0.5 * ph);
    }
    if (flags & 3) {
        float rate = (float)((double)(c->mode == 1 ? 250000.0 : 240000.0) * c->audio_up / c->audio_decim);
        double k = 2.0 * (double)rate * (double)((flags & 1) ? 75.0f : 50.0f) * 1e-6;
        c->de_b = 1.0 / (1.0 + k);
        c->de_a1 = (1.0 - k) / (1.0 + k);
    }
}

int orc_chain_audio_per_block(const orc_chain *c) { return c->n_audio; }
void orc_chain_set_paths(orc_chain *c, int mask) { c->paths = mask; }

/* src/fm_radio.cpp:290-298: NaN -> 0, else static_cast<short>(x*16384*mult) — x86 cvttss2si then the low 16 bits */
static int16_t quantise(float v, int mult) {
    if (isnan(v)) return 0;
    float s = (v * 16384.0f) * (float)mult;
    int32_t w = (s > -2147483648.0f && s < 2147483648.0f) ? (int32_t)s : INT32_MIN;
    return (int16_t)(uint16_t)((uint32_t)w & 0xFFFFu);
}

int orc_chain_block(orc_chain *c, const uint8_t *raw, int16_t *audio) {
    /* ---- rf_thread: :66-84 ---- */
    orc_unpack(raw, ORC_BLOCK_BYTES, c->iq);
    for (int i = 0; i < NIQ; i++) { c->I[i] = c->iq[2 * i]; c->Q[i] = c->iq[2 * i + 1]; }
    orc_fir_decim_iq(c->If, c->Qf, c->I, c->Q, NIQ, c->h_rf, NT, c->zi_i, c->zi_q, 10);
    orc_demod(c->If, c->Qf, NIF, c->demod);

    /* ---- mono_stero_thread: :226-307 ---- */
    if (c->paths & 1) {
        int na = c->n_audio;
        if (c->mode != 0) orc_resample(c->mono, 0, c->demod, NIF, c->h_mono, c->audio_taps, c->zi_mono, c->nzi_a, c->audio_decim, c->audio_up, 0); /* :228 */
        else orc_fir_decim(c->mono, c->demod, NIF, c->h_mono, NT, c->zi_mono, c->nzi_a, 5);                                                   /* :258 */
        /* Q7: `mixed.clear()` (:307) kills the stereo path from block 1 on in the shipped binary */
        int stereo_live = (c->profile == ORC_PROFILE_INTENT) || c->block_id == 0;
        if (stereo_live) {
            orc_fir_decim(c->pilot, c->demod, NIF, c->h_pilot, NT, c->zi_pilot, c->nzi_a, 1);                 /* :232/:261 */
            orc_pll(c->nco, c->pilot, NIF, 19e3f, 240e3f, 2.0f, 0.0f, 0.01f, c->pll_st);                       /* :233/:262 */
            orc_fir_decim(c->sbpf, c->demod, NIF, c->h_sbpf, NT, c->zi_sbpf, c->nzi_a, 1);                    /* :236/:265 */
            for (int i = 0; i < NIF; i++) c->mixed[i] = c->sbpf[i] * c->nco[i];                                /* :240-243/:269-272 */
            if (c->mode == 1) orc_resample(c->stereo, na, c->mixed, NIF, c->h_stereo, c->audio_taps, c->zi_stereo, c->nzi_a, 5, c->audio_up, 0); /* :245 (Q14) */
            else if (c->mode == 2) orc_resample(c->stereo, 0, c->mixed, NIF, c->h_stereo, c->audio_taps, c->zi_stereo, c->nzi_a, c->audio_decim, c->audio_up, 0); /* the call :245 meant to make */
            else orc_fir_decim(c->stereo, c->mixed, NIF, c->h_stereo, NT, c->zi_stereo, c->nzi_a, 5);                                          /* :274 */
        } else {
            for (int i = 0; i < na; i++) c->stereo[i] = 0.0f;
        }
        for (int i = 0; i < na; i++) { /* :247-252/:277-282 */
            float m = c->mono[i];
            if (c->mono_delay) m = i >= c->mono_delay ? c->mono[i - c->mono_delay] : c->mono_tail[i]; /* quality: L+R as late as L-R */
            float l = (m + c->stereo[i]) / 2.0f, r = (m - c->stereo[i]) / 2.0f;
            if (c->de_b != 0.0) { /* quality: 1/(1 + s tau), bilinear, in double (scipy.signal.lfilter's recurrence); a NaN sample counts as 0 */
                float in[2] = {isnan(l) ? 0.0f : l, isnan(r) ? 0.0f : r};
                float out[2];
                for (int ch = 0; ch < 2; ch++) {
                    double y = c->de_b * ((double)in[ch] + c->de_state[2 * ch]) - c->de_a1 * c->de_state[2 * ch + 1];
                    c->de_state[2 * ch] = (double)in[ch];
                    c->de_state[2 * ch + 1] = y;
                    out[ch] = (float)y;
                }
                l = out[0]; r = out[1];
            }
            c->audio_f[2 * i] = l; c->audio_f[2 * i + 1] = r;
            if (audio) { audio[2 * i] = quantise(l, c->mult); audio[2 * i + 1] = quantise(r, c->mult); }
        }
        for (int j = 0; j < c->mono_delay; j++) c->mono_tail[j] = c->mono[na - c->mono_delay + j];
    }

    /* ---- rds_thread (:395-411) + frame_thread: not in mode 1 (:324, :446); mode 2 has mode 0's 240 kHz IF, so it runs ---- */
    c->nbits = 0; c->nev = 0;
    if (c->mode != 1 && (c->paths & 2)) {
        orc_fir_decim(c->rbpf, c->demod, NIF, c->h_rbpf, NT, c->zi_rbpf, NT - 1, 1);
        orc_pll_combine(c->rsq, c->rnco, c->rbpf, NIF, c->h_sq, NT, c->zi_sq, 114000.0f, 240000.0f, 0.5f, c->rds_phase, 0.001f, c->rds_pll_st);
        orc_fir_mixer(c->rlpf, c->rnco, c->rbpf, NIF, c->h_lpf3k, NT, c->zi_lpf);
        c->rlpf[NIF] = 0.0f; /* the reference's element 15360 is never consumed (Q8) */
        orc_resample(c->rres, 0, c->rlpf, NIF + 1, c->h_anti, NT * 19, c->zi_anti, NT * 19 - 1, 80, 19, 1);
        orc_fir_decim(c->rrrc, c->rres, c->n_rds, c->h_rrc, NT, c->zi_rrc, NT - 1, 1);
        c->nbits = orc_rds_decode_block(c->dec, c->rrrc, c->n_rds, c->bits, c->ev, 256, &c->nev);
    }
    c->block_id++;
    return 2 * c->n_audio;
}

const float *orc_chain_tap(const orc_chain *c, int which, int *n) {
    const float *p = NULL; int len = 0;
    switch (which) {
        case ORC_TAP_DEMOD: p = c->demod; len = NIF; break;
        case ORC_TAP_MONO: p = c->mono; len = c->n_audio; break;
        case ORC_TAP_PILOT: p = c->pilot; len = NIF; break;
        case ORC_TAP_NCO: p = c->nco; len = NIF; break;
        case ORC_TAP_STEREO_BPF: p = c->sbpf; len = NIF; break;
        case ORC_TAP_STEREO: p = c->stereo; len = c->n_audio; break;
        case ORC_TAP_RDS_BPF: p = c->rbpf; len = NIF; break;
        case ORC_TAP_RDS_SQ: p = c->rsq; len = NIF; break;
        case ORC_TAP_RDS_NCO: p = c->rnco; len = NIF + 1; break;
        case ORC_TAP_RDS_LPF: p = c->rlpf; len = NIF; break;
        case ORC_TAP_RDS_RES: p = c->rres; len = c->n_rds; break;
        case ORC_TAP_RDS_RRC: p = c->rrrc; len = c->n_rds; break;
        case ORC_TAP_I: p = c->If; len = NIF; break;
        case ORC_TAP_Q: p = c->Qf; len = NIF; break;
        case ORC_TAP_AUDIO_F: p = c->audio_f; len = 2 * c->n_audio; break;
        default: break;
    }
    if (n) *n = len;
    return p;
}

int orc_chain_rds_bits(const orc_chain *c, uint8_t *bits, int cap) {
    int n = c->nbits < cap ? c->nbits : cap;
    if (bits) memcpy(bits, c->bits, (size_t)n);
    return c->nbits;
}
int orc_chain_rds_events(const orc_chain *c, orc_rds_event *ev, int cap) {
    int n = c->nev < cap ? c->nev : cap;
    if (ev) memcpy(ev, c->ev, sizeof(orc_rds_event) * (size_t)n);
    return c->nev;
}
int orc_chain_rds_offset(const orc_chain *c) { return orc_rds_initial_offset(c->dec); }
