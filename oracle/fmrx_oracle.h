/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference FM receive chain.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library, and
 * only as the checker.  The product (libfmrx.so) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against the unmodified reference objects
 * (oracle/_ref/libfmref.so, built by oracle/Makefile from /root/reference/src) and against the reference binary's
 * stdout/stderr in tests/test_oracle_vs_reference.py, and against the committed fixtures in tests/golden/ (generated
 * from the reference by tests/golden/make_golden.py) in tests/test_oracle_golden.py.
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef FMRX_ORACLE_H
#define FMRX_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- filter design: src/filter.cpp:19-38, 41-60, 63-93 ------------------------------------------------------- */
void orc_design_lpf(float Fs, float Fc, unsigned short ntaps, float *h);
void orc_design_bpf(float Fb, float Fe, float Fs, int ntaps, float *h);
void orc_design_rrc(float Fs, int ntaps, float *h);

/* ---- a1: src/iofunc.cpp:61-69 -------------------------------------------------------------------------------- */
void orc_unpack(const uint8_t *raw, size_t n, float *dst);

/* ---- a6/a8/a9: src/filter.cpp:126-219.  y is OVERWRITTEN (fresh-vector semantics); nzi may exceed ntaps-1 (mode 1
 * sizes every state from the 3624-tap filter, src/fm_radio.cpp:189-193) -------------------------------------- */
int orc_fir_decim(float *y, const float *x, int n, const float *h, int ntaps, float *zi, int nzi, int decim);
int orc_fir_decim_iq(float *yi, float *yq, const float *xi, const float *xq, int n, const float *h, int ntaps,
                     float *zii, float *ziq, int decim);
/* ---- a10/a11/a12: src/filter.cpp:222-339.  gain_up!=0 multiplies by `up` (the RDS variant, :333); only the first
 * ny_limit outputs are produced (<=0: all floor(n*up/decim)) -------------------------------------------------- */
int orc_resample(float *y, int ny_limit, const float *x, int n, const float *h, int ntaps, float *zi, int nzi,
                 int decim, int up, int gain_up);
/* ---- a15: src/filter.cpp:373-401.  x = NCO (n+1 long in the reference; only n are meaningful), x1 = signal (n) - */
int orc_fir_mixer(float *y, const float *nco, const float *sig, int n, const float *h, int ntaps, float *zi);
/* ---- a7: src/rf_module.cpp:13-34 ----------------------------------------------------------------------------- */
void orc_demod(const float *I, const float *Q, int n, float *dst);
/* ---- a13: src/helper.cpp:13-57; st = {integrator, phaseEst, feedbackI, feedbackQ, trigOffset, ncoLast} -------- */
void orc_pll(float *nco, const float *x, int n, float freq, float Fs, float scale, float phase_adj, float bw, float *st);
/* ---- a14: src/helper.cpp:108-173; nco gets n+1 elements ------------------------------------------------------ */
void orc_pll_combine(float *y, float *nco, const float *x, int n, const float *h, int ntaps, float *zi, float freq,
                     float Fs, float scale, float phase_adj, float bw, float *st);

/* ---- a17-a20: src/fm_radio.cpp:444-729 ------------------------------------------------------------------------ */
enum { ORC_EV_GOOD = 0, ORC_EV_FALSE = 1, ORC_EV_RESYNC = 2 };
typedef struct {
    int32_t block;    /* block id in which the event was printed */
    int32_t kind;     /* ORC_EV_* */
    int32_t letter;   /* 0..3 = A..D, -1 for resync */
    uint32_t position; /* `printposition` */
} orc_rds_event;

typedef struct orc_rds_decoder orc_rds_decoder;
orc_rds_decoder *orc_rds_decoder_create(void);
void orc_rds_decoder_destroy(orc_rds_decoder *);
/* one block of RRC output (3648 samples).  bits_out (may be NULL) receives this block's differentially decoded
 * bits (75 in block 0, 76 after); events are appended to ev[0..ev_cap).  Returns #bits; *n_ev = #events. */
int orc_rds_decode_block(orc_rds_decoder *, const float *rrc, int n, uint8_t *bits_out, orc_rds_event *ev, int ev_cap,
                         int *n_ev);
int orc_rds_initial_offset(const orc_rds_decoder *);
int orc_rds_start_pos(const orc_rds_decoder *);
/* renders exactly the stderr lines frame_thread prints for one block (offset line in block 0, banner, events) */
int orc_rds_format_block(int block_id, int initial_offset, const orc_rds_event *ev, int n_ev, char *buf, int cap);

/* ---- the chain: thread bodies of src/fm_radio.cpp:31-441 composed sequentially ------------------------------- */
enum { ORC_PROFILE_BINARY = 0, ORC_PROFILE_INTENT = 1 };
enum {
    ORC_TAP_DEMOD = 0, ORC_TAP_MONO, ORC_TAP_PILOT, ORC_TAP_NCO, ORC_TAP_STEREO_BPF, ORC_TAP_STEREO,
    ORC_TAP_RDS_BPF, ORC_TAP_RDS_SQ, ORC_TAP_RDS_NCO, ORC_TAP_RDS_LPF, ORC_TAP_RDS_RES, ORC_TAP_RDS_RRC,
    ORC_TAP_I, ORC_TAP_Q, ORC_TAP_AUDIO_F, ORC_TAP_COUNT
};
#define ORC_BLOCK_BYTES 307200
typedef struct orc_chain orc_chain;
orc_chain *orc_chain_create(int mode, int profile);
void orc_chain_destroy(orc_chain *);
int orc_chain_audio_per_block(const orc_chain *); /* 3072 (mode 0), 2949 (mode 1), 2822 (mode 2: 44.1 kHz, not in the reference) */
/* one 307200-byte block -> 2*audio_per_block int16 (L,R interleaved).  Returns #int16 written. */
int orc_chain_block(orc_chain *, const uint8_t *iq, int16_t *audio);
const float *orc_chain_tap(const orc_chain *, int which, int *n);
/* RDS results of the most recent block */
int orc_chain_rds_bits(const orc_chain *, uint8_t *bits, int cap);
int orc_chain_rds_events(const orc_chain *, orc_rds_event *ev, int cap);
int orc_chain_rds_offset(const orc_chain *);
/* stage switches for timing: bit0 mono/stereo, bit1 rds (default 3) */
void orc_chain_set_paths(orc_chain *, int mask);
/* FMRX_QUALITY_* of include/fmrx.h: the quality profile is NOT in the reference; call right after create */
void orc_chain_set_quality(orc_chain *, int flags);

#ifdef __cplusplus
}
#endif
#endif
