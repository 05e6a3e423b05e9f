/* fmrx — B200-native FM broadcast receive chain.  C-ABI of libfmrx.so.
 *
 * The reference (m1nty/Real-Time-Software-Defined-Radio) has no FFI: its only interfaces are the `fm_radio` process
 * contract and the C++ free functions of src/filter.h, src/helper.h, src/rf_module.h and src/iofunc.h.  Every entry
 * point below names the reference function (file:line under /root/reference) whose arithmetic it reproduces on the
 * GPU; include/fmrx_dropin.hpp re-exposes them under the reference's own C++ names and signatures.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are HOST pointers unless the name says `_device`;
 *  - every call returns FMRX_OK or a negative fmrx_status; nothing ever calls exit(); there is NO CPU fallback — with
 *    no usable CUDA device every compute entry returns FMRX_ERR_CUDA (fmrx_last_error() has the driver's message);
 *  - batched entries process `n_streams` independent streams x `n_blocks` consecutive blocks per stream in one
 *    launch; arrays are stream-major ([stream][block][sample]); per-stream filter state (`zi`, PLL state) is carried
 *    from block to block inside the call and written back at the end, exactly as if the reference function had been
 *    called once per block (the state is saved one sample late, and that is reproduced);
 *  - `exact != 0` selects the reference's rounding: separate fp32 multiply and add per tap, taps in ascending order
 *    (results are bit-identical to the reference built with g++ -O3 on x86-64); `exact == 0` uses fused multiply-add
 *    (<= 1e-6 relative RMS from the reference).
 */
#ifndef FMRX_H
#define FMRX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FMRX_VERSION 101
#define FMRX_BLOCK_BYTES 307200 /* src/fm_radio.cpp:23 */
#define FMRX_IF_PER_BLOCK 15360  /* 307200 / 2 / 10 */
#define FMRX_RDS_PER_BLOCK 3648  /* floor(15361*19/80), src/filter.cpp:304 */
#define FMRX_MAX_TAPS 151        /* register-tiled FIR kernels are specialised for the reference's 151 taps */
/* per stream per block: the sliding syndrome window of frame_thread takes at most 77 positions in a block (27 carried bits +
 * 76 new ones, src/fm_radio.cpp:631-713) and each can print one syndrome line and one re-sync line (:649-704): 154 */
#define FMRX_MAX_EVENTS 160
#define FMRX_MAX_BITS 80         /* per stream per block */

typedef enum {
    FMRX_OK = 0,
    FMRX_ERR_ARG = -1,    /* bad argument (null pointer, unsupported size) */
    FMRX_ERR_CUDA = -2,   /* CUDA runtime / driver error, or no device */
    FMRX_ERR_ALLOC = -3,  /* device or pinned-host allocation failed */
    FMRX_ERR_STATE = -4,  /* handle used in the wrong state */
    FMRX_ERR_TIMEOUT = -5, /* fmrx_ring_acquire / fmrx_ring_next: nothing became available in time */
    FMRX_ERR_EOF = -6     /* fmrx_ring_next after fmrx_ring_close: every committed step has been delivered */
} fmrx_status;

const char *fmrx_last_error(void);
int fmrx_version(void);
int fmrx_device_count(void);

/* ---- host-side filter design (bit-identical taps) ----------------------------------------------------------- */
int fmrx_design_lpf(float Fs, float Fc, unsigned short ntaps, float *h); /* impulseResponseLPF, src/filter.cpp:19-38 */
int fmrx_design_bpf(float Fb, float Fe, float Fs, int ntaps, float *h);  /* impulseResponseBPF, src/filter.cpp:41-60 */
int fmrx_design_rrc(float Fs, int ntaps, float *h);                      /* impulseResponseRRC, src/filter.cpp:63-93 */

/* quality-profile design helpers (host, double): response of a real FIR at f; impulseResponseBPF scaled to unit gain at its band
 * centre; the NCO phase adjust that lines the regenerated RDS carrier up with the RDS band (-arg H_sq(f2) / 2); the bilinear
 * de-emphasis coefficients of y[n] = b (x[n] + x[n-1]) - a1 y[n-1] */
int fmrx_fir_response(const float *h, int ntaps, float Fs, float f, double *mag, double *phase);
int fmrx_design_bpf_unity(float Fb, float Fe, float Fs, int ntaps, float *h);
int fmrx_rds_auto_phase(const float *h_sq, int ntaps, float Fs, float f2, float *phase_adj);
int fmrx_deemphasis_coeffs(float tau_us, float Fs, double *b, double *a1);

/* ---- function-level operators (GPU; host buffers) ------------------------------------------------------------ */
/* de-emphasis + the reference's quantiser (src/fm_radio.cpp:290-298) on interleaved L,R float audio: audio_f:[S][B][2n] filtered in
 * place, audio:[S][B][2n] int16 (may be NULL), state:[S][4] = x[-1], y[-1] of L, then of R (zero before the first block) */
int fmrx_deemphasis(float *audio_f, int16_t *audio, int n_streams, int n_blocks, int n, float tau_us, float Fs, int mult, float *state);
/* readStdInBlock's conversion, src/iofunc.cpp:61-69: out[k] = (raw[k]-128)/128 */
int fmrx_unpack_iq(const uint8_t *raw, size_t n, float *out);
/* convolveWithDecim / convolveWithDecimPointer, src/filter.cpp:126-185.  x:[S][B][n] y:[S][B][n/decim] zi:[S][nzi];
 * ntaps must be 151, decim in {1,5,10}; nzi >= 150 (the last 150 entries are the live state, src/fm_radio.cpp:189-193) */
int fmrx_fir_decim(float *y, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps, float *zi,
                   int nzi, int decim, int exact);
/* convolveWithDecimIQ, src/filter.cpp:187-219 */
int fmrx_fir_decim_iq(float *yi, float *yq, const float *xi, const float *xq, int n_streams, int n_blocks, int n,
                      const float *h, int ntaps, float *zii, float *ziq, int decim, int exact);
/* convolveWithDecimMode1 / ...Pointer / ...RDS, src/filter.cpp:222-339.  y:[S][B][ny] with ny = ny_limit>0 ?
 * min(ny_limit, n*up/decim) : n*up/decim; gain_up multiplies by `up` (:333).  Any ntaps. */
int fmrx_resample(float *y, int ny_limit, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps,
                  float *zi, int nzi, int decim, int up, int gain_up, int exact);
/* convolveWithDecimAndMixer at its call site src/fm_radio.cpp:404 (src/filter.cpp:373-401): nco:[S][B][n] (element k
 * pairs with sig element k), sig:[S][B][n], y:[S][B][n]; half-weight history (Q8) */
int fmrx_fir_mixer(float *y, const float *nco, const float *sig, int n_streams, int n_blocks, int n, const float *h,
                   int ntaps, float *zi);
/* fmDemodArctan, src/rf_module.cpp:13-34 (previous sample reset at every block start, Q3) */
int fmrx_demod(const float *I, const float *Q, int n_streams, int n_blocks, int n, float *out);
/* fmPLL, src/helper.cpp:13-57.  state:[S][6] = {integrator, phaseEst, feedbackI, feedbackQ, trigOffset, ncoLast} */
int fmrx_pll(float *nco, const float *x, int n_streams, int n_blocks, int n, float freq, float Fs, float scale,
             float phase_adj, float bw, float *state);
/* pllCombine, src/helper.cpp:108-173: y = BPF(x^2) (zi holds squares), nco as fmrx_pll on y.  The reference's
 * untrimmed (n+1)-th NCO element is returned in state[5] (ncoLast). */
int fmrx_pll_combine(float *y, float *nco, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps,
                     float *zi, float freq, float Fs, float scale, float phase_adj, float bw, float *state);
/* rf_thread's per-block work fused, src/fm_radio.cpp:66-84: u8 IQ -> unpack -> deinterleave -> 151-tap LPF, /decim on
 * I and Q -> discriminator.  raw:[S][B][2*n] bytes, demod:[S][B][n/decim]; optional yi,yq (may be NULL) receive the
 * filtered I/Q.  Always reference-exact. */
int fmrx_frontend(float *demod, float *yi, float *yq, const uint8_t *raw, int n_streams, int n_blocks, int n,
                  const float *h, int ntaps, float *zii, float *ziq, int decim);

/* ---- RDS clock/data recovery + frame sync: frame_thread, src/fm_radio.cpp:444-729 --------------------------- */
enum { FMRX_EV_GOOD = 0, FMRX_EV_FALSE = 1, FMRX_EV_RESYNC = 2 };
typedef struct {
    int32_t block;     /* per-stream block id the event belongs to */
    int32_t kind;      /* FMRX_EV_* */
    int32_t letter;    /* 0..3 = offset word A..D; -1 for a resync */
    uint32_t position; /* the reference's `printposition` */
} fmrx_rds_event;

/* opaque per-stream decoder state (block counter, sampling phase, Manchester alignment, carried bits, sync counters) */
#define FMRX_RDS_STATE_WORDS 160
/* rrc:[S][B][n] (n = 3648 in the chain; n/24 <= 160).  bits:[S][B][FMRX_MAX_BITS] n_bits:[S][B];
 * events:[S][B][FMRX_MAX_EVENTS] n_events:[S][B]; state:[S][FMRX_RDS_STATE_WORDS] int32, zero-initialised by the
 * caller before the first block. */
int fmrx_rds_decode(const float *rrc, int n_streams, int n_blocks, int n, uint8_t *bits, int32_t *n_bits,
                    fmrx_rds_event *events, int32_t *n_events, int32_t *state);
/* sampling phase picked in block 0 (`initial_offset`, :503-517) from a decoder state */
int fmrx_rds_state_offset(const int32_t *state);
/* renders the exact stderr lines frame_thread prints for one block; returns the length written (excluding NUL) */
int fmrx_rds_format_block(int block_id, int initial_offset, const fmrx_rds_event *ev, int n_ev, char *buf, int cap);

/* ---- the batched receive chain ------------------------------------------------------------------------------ */
enum { FMRX_PROFILE_BINARY = 0, FMRX_PROFILE_INTENT = 1 };   /* SURVEY App. A */
/* FMRX_PATH_RDS_STAGES (with FMRX_PATH_RDS): run the RDS back end stage by stage at full rate, as the reference does --
 * mixer + 3 kHz LPF, 19/80 resampler, RRC -- so that every intermediate signal exists (FMRX_TAP_RDS_LPF / _RES / _RRC).
 * Without it (the default) those three filters run as one composite polyphase filter evaluated only at the 152 samples
 * per block the decoder reads (csrc/fmrx_rdsfast.cu); the RRC tap then holds just those samples. */
enum { FMRX_PATH_AUDIO = 1, FMRX_PATH_RDS = 2, FMRX_PATH_RDS_STAGES = 4 };
/* FMRX_NUMERICS_REFERENCE (default): every filter of the audio path and every filter ahead of the 19 kHz loop with the
 *   reference's two roundings per tap (bit-identical float audio); on the RDS branch the 54-60 kHz band-pass likewise,
 *   the squared-input filter of pllCombine and everything behind the 114 kHz loop fused-multiply-add.
 * FMRX_NUMERICS_STRICT: additionally pllCombine's filter with the reference's double products accumulated into a float
 *   sum (src/helper.cpp:139), so the 114 kHz loop, its NCO and the mixer product are bit-identical to the reference's;
 *   with FMRX_PATH_RDS_STAGES the mixer filter, the 19/80 resampler and the RRC keep two roundings per tap as well and the
 *   whole RDS branch is bit-identical (decoded bits then equal the reference's by construction, whatever the input).
 * FMRX_NUMERICS_FMA: fused multiply-add wherever the stated tolerance (1e-5 relative RMS on float audio) allows: the mono
 *   and stereo low-pass, the 22-54 kHz band-pass and the whole RDS branch; only the pilot band-pass stays exact (anything
 *   ahead of a PLL decides the fp32 rounding of the oscillator argument, src/helper.cpp:41). */
enum { FMRX_NUMERICS_REFERENCE = 0, FMRX_NUMERICS_FMA = 1, FMRX_NUMERICS_STRICT = 2 };

/* `quality` profile: what the reference's own report proposes and the program never got -- NEVER the default,
 * because with any of these set the output is no longer the reference's:
 *  FMRX_QUALITY_DEEMPH_75 / _50: de-emphasis 1 / (1 + s tau), tau = 75 us (Americas, Korea) or 50 us, bilinear transform at the
 *    audio rate, on L and R ahead of the quantiser (the reference quantises the low-pass output as is, src/fm_radio.cpp:277-299);
 *  FMRX_QUALITY_UNITY_BPF: the four band-pass filters scaled to unit gain at their band centre (impulseResponseBPF has 0.308 at
 *    19 kHz, src/filter.cpp:41-60), the stereo mixer with the x2 a product of two cosines needs (the reference mixes x1,
 *    src/fm_radio.cpp:271; its Python model x2, model/fmMonoBlock.py:155-156), and the L+R branch delayed by the 75 IF samples
 *    (15 at 48 kHz) the L-R branch spends in its extra band-pass -- together: L-R at the gain and timing of L+R, i.e. stereo
 *    separation instead of the reference's level-and-delay mismatch;
 *  FMRX_QUALITY_AUTO_RDS_PHASE: the 114 kHz NCO's phase adjust computed from the response of pllCombine's filter at 114 kHz
 *    (fmrx_rds_auto_phase) instead of the hand-tuned constant of src/fm_radio.cpp:342,400 (14 degrees off on the reference's taps). */
enum { FMRX_QUALITY_DEEMPH_75 = 1, FMRX_QUALITY_DEEMPH_50 = 2, FMRX_QUALITY_UNITY_BPF = 4, FMRX_QUALITY_AUTO_RDS_PHASE = 8 };

typedef struct {
    int32_t mode;       /* 0: 2.4 Msps, /10, /5, +RDS ; 1: 2.5 Msps, /10, x24 /125, no RDS (src/fm_radio.cpp:36-37,174-180);
                         * 2 (extension, not in the reference's main()): mode 0's front end and RDS path with the audio
                         * resampled x147 /800 to 44.1 kHz by the reference's polyphase resampler (BASELINE config 2) */
    int32_t profile;    /* FMRX_PROFILE_* */
    int32_t n_streams;  /* independent stations */
    int32_t max_blocks; /* largest n_blocks a process call will be given */
    int32_t device;     /* CUDA device ordinal */
    int32_t paths;      /* FMRX_PATH_* mask; 0 = all the mode has */
    int32_t numerics;   /* FMRX_NUMERICS_* */
    int32_t quality;    /* FMRX_QUALITY_* mask; 0 = the reference's receiver */
} fmrx_config;

typedef struct fmrx_batch fmrx_batch;

/* outputs of one process call; any pointer may be NULL to skip that copy.  Host pointers for fmrx_batch_process,
 * device pointers for fmrx_batch_process_device. */
typedef struct {
    int16_t *audio;          /* [S][B][2*audio_per_block] interleaved L,R (src/fm_radio.cpp:286-302) */
    float *audio_f;          /* same shape, before quantisation */
    uint8_t *rds_bits;       /* [S][B][FMRX_MAX_BITS] */
    int32_t *rds_n_bits;     /* [S][B] */
    fmrx_rds_event *rds_events; /* [S][B][FMRX_MAX_EVENTS] */
    int32_t *rds_n_events;   /* [S][B] */
} fmrx_outputs;

int fmrx_batch_create(const fmrx_config *cfg, fmrx_batch **out);
void fmrx_batch_destroy(fmrx_batch *);
int fmrx_batch_audio_per_block(const fmrx_batch *); /* 3072 (mode 0) / 2949 (mode 1) / 2822 (mode 2) */
int fmrx_batch_reset(fmrx_batch *);                 /* back to block 0 with the reference's initial state */
/* host -> host.  iq:[S][n_blocks][307200] bytes.  Copies are staged through pinned rings and overlapped with the
 * kernels on separate CUDA streams (the replacement for the reference's producer/consumer threads). */
int fmrx_batch_process(fmrx_batch *, const uint8_t *iq, int n_blocks, const fmrx_outputs *out);
/* host -> host, asynchronous: enqueues the H2D copy of the whole step, the three-phase pipeline and the D2H copy of
 * the results, and returns a ticket; consecutive submits overlap (step k+1's ingest runs under step k's kernels, which
 * is what replaces the reference's rf_thread -> queue -> consumer threads, src/fm_radio.cpp:86-138).  `iq` and the
 * output buffers must be page-locked (fmrx_pinned_alloc) and stay untouched until fmrx_batch_wait(ticket) returns. */
int fmrx_batch_submit(fmrx_batch *, const uint8_t *iq, int n_blocks, const fmrx_outputs *out, long long *ticket);
int fmrx_batch_wait(fmrx_batch *, long long ticket);
/* returns when the host -> device copy of that step has finished (its `iq` buffer may be refilled; the results are not there yet).
 * For callers that arbitrate the host link between several handles / processes: on hosts that feed four GPUs at once faster than
 * eight, letting two halves of the ranks take turns on the link moves more than letting all of them copy at once. */
int fmrx_batch_wait_ingest(fmrx_batch *, long long ticket);
/* device -> device, asynchronous on the handle's own (non-blocking) streams; fmrx_batch_sync() waits.  The handle's
 * streams do not synchronise with any stream of the caller: iq_device must be complete before the call (synchronise the
 * stream that produced it) and the outputs must not be read before fmrx_batch_sync() or a wait on fmrx_batch_cuda_stream(). */
int fmrx_batch_process_device(fmrx_batch *, const uint8_t *iq_device, int n_blocks, const fmrx_outputs *out_device);
int fmrx_batch_sync(fmrx_batch *);
void *fmrx_batch_cuda_stream(fmrx_batch *);         /* cudaStream_t on which a call's outputs are completed */
/* the device-resident path is a three-stage pipeline over three streams: phase 0 = front end + filters ahead of the
 * PLLs, 1 = the PLLs, 2 = everything after them (+ output copies).  Consecutive fmrx_batch_process_device calls overlap:
 * call k's PLLs run beside call k+1's phase 0 and call k-1's phase 2. */
void *fmrx_batch_cuda_stream_phase(fmrx_batch *, int phase);
/* SMs the PLL phase / the filter phases of the device-resident pipeline own (green-context partition; both 0 when the
 * phases share the whole device: small batches, FMRX_PLL_SMS=0, or a driver without green contexts) */
int fmrx_batch_partition(const fmrx_batch *, int *pll_sms, int *filter_sms);
long long fmrx_batch_launch_count(const fmrx_batch *); /* kernels launched by this handle so far */
/* per-stream initial_offset of the RDS decoder (host int32[S]) */
int fmrx_batch_rds_offsets(fmrx_batch *, int32_t *offsets);
/* the 114 kHz NCO's phase adjust in use (the reference's constant, or the computed one under FMRX_QUALITY_AUTO_RDS_PHASE); the
 * setter is for experiments (phase sweeps): it takes effect from the next process call */
float fmrx_batch_rds_phase(const fmrx_batch *);
int fmrx_batch_set_rds_phase(fmrx_batch *, float phase_adj);

/* intermediate signals of the most recent process call, for per-stage parity tests: copies [S][n_blocks][len] floats */
enum {
    FMRX_TAP_DEMOD = 0, FMRX_TAP_MONO, FMRX_TAP_PILOT, FMRX_TAP_NCO, FMRX_TAP_STEREO_BPF, FMRX_TAP_STEREO,
    FMRX_TAP_RDS_BPF, FMRX_TAP_RDS_SQ, FMRX_TAP_RDS_NCO, FMRX_TAP_RDS_LPF, FMRX_TAP_RDS_RES, FMRX_TAP_RDS_RRC,
    FMRX_TAP_COUNT
};
int fmrx_batch_tap_len(const fmrx_batch *, int which); /* samples per block of that signal */
int fmrx_batch_tap(fmrx_batch *, int which, float *dst);

/* per-stage device timing inside the real chain: while enabled, every stage of every enqueued chain is bracketed by
 * CUDA events on the stream it runs on, and the device-resident path runs its three phases back to back on ONE stream
 * (no overlap), so that each stage is timed alone; fmrx_batch_stage_times() synchronises and returns the accumulated
 * milliseconds and bracket counts per stage since profiling was (re-)enabled. */
enum {
    FMRX_STAGE_FRONTEND = 0, FMRX_STAGE_MONO, FMRX_STAGE_PILOT_BPF, FMRX_STAGE_STEREO_BPF, FMRX_STAGE_RDS_BPF,
    FMRX_STAGE_RDS_SQ_BPF, FMRX_STAGE_PLL, FMRX_STAGE_STEREO_LPF, FMRX_STAGE_COMBINE, FMRX_STAGE_RDS_MIX_LPF,
    FMRX_STAGE_RDS_RESAMPLE, FMRX_STAGE_RDS_RRC, FMRX_STAGE_RDS_DECODE, FMRX_STAGE_RDS_SYMBOLS,
    FMRX_STAGE_BPF_FUSED, /* the band-pass filters of the discriminator output that share one launch (csrc/fmrx_fir.cu fir151_multi_kernel) */
    FMRX_STAGE_COUNT
};
int fmrx_batch_profile(fmrx_batch *, int enable); /* 0 off, 1 serialised per-stage timing, 2 timeline: keep the pipeline */
/* start / end of every bracket recorded since profiling was enabled, in ms after the first bracket's start; returns the
 * number of brackets written (<= cap) or a negative status.  Call before fmrx_batch_stage_times (which consumes them). */
int fmrx_batch_timeline(fmrx_batch *, int cap, int32_t *stage, float *t0_ms, float *t1_ms);
int fmrx_batch_stage_times(fmrx_batch *, double *ms /*[FMRX_STAGE_COUNT]*/, long long *count /*[FMRX_STAGE_COUNT] or NULL*/);

/* opaque state blob (all filter histories, PLL states, decoder states, block counter) for checkpoint / resume.  The blob
 * starts with a header (magic, layout version, mode, profile, n_streams, paths, size); set_state returns FMRX_ERR_ARG for a
 * blob that is truncated, was written by another layout version or comes from a handle of another shape. */
size_t fmrx_batch_state_bytes(const fmrx_batch *);
int fmrx_batch_get_state(fmrx_batch *, void *blob, size_t bytes);
int fmrx_batch_set_state(fmrx_batch *, const void *blob, size_t bytes);
long long fmrx_batch_block_id(const fmrx_batch *); /* blocks consumed per stream so far */

/* ---- ingest / egress ring ------------------------------------------------------------------
 * A bounded ring of page-locked host slots in front of a batch handle: what replaces rf_thread -> queue -> consumer
 * threads (src/fm_radio.cpp:86-138: five-slot ring, condition variables) when one process feeds many stations.
 * ONE producer thread acquires a free slot (blocking while every slot is in flight: that is the back-pressure the
 * reference gets from its bounded queues), fills iq[S][n_blocks][307200] and commits it, which enqueues the H2D copy,
 * the three-phase pipeline and the D2H copy of the results; ONE consumer thread (the same or another) takes the steps in
 * order with fmrx_ring_next -- it blocks until that step's results are in the slot's host buffers -- and releases the
 * slot for reuse.  While a ring is attached no other process / submit call may be made on the handle. */
typedef struct fmrx_ring fmrx_ring;
int fmrx_ring_create(fmrx_batch *, int n_slots, int n_blocks, fmrx_ring **out);
void fmrx_ring_destroy(fmrx_ring *);
int fmrx_ring_acquire(fmrx_ring *, int timeout_ms, uint8_t **iq);   /* producer; timeout_ms < 0: wait for ever; FMRX_ERR_TIMEOUT */
int fmrx_ring_commit(fmrx_ring *);                                   /* producer: submit the acquired slot (n_blocks blocks per station) */
/* the same for a slot that holds fewer blocks -- the short last step at end of input; arrays of that step are [S][n][..] */
int fmrx_ring_commit_blocks(fmrx_ring *, int n_blocks);
int fmrx_ring_close(fmrx_ring *);                                    /* producer: end of input */
/* consumer: oldest committed step.  `out` receives pointers into the slot's pinned buffers (audio, rds_bits, rds_n_bits,
 * rds_events, rds_n_events; audio_f is not carried), valid until fmrx_ring_release.  FMRX_ERR_TIMEOUT / FMRX_ERR_EOF. */
int fmrx_ring_next(fmrx_ring *, int timeout_ms, fmrx_outputs *out);
int fmrx_ring_release(fmrx_ring *);
int fmrx_ring_step_blocks(fmrx_ring *);                              /* consumer: blocks per station of the step taken and not yet released */
int fmrx_ring_in_flight(fmrx_ring *);                                /* committed and not yet released */

/* page-locked host memory for the ingest / egress rings (fmrx_batch_process copies asynchronously only from/to it) */
int fmrx_pinned_alloc(void **ptr, size_t bytes);
int fmrx_pinned_free(void *ptr);

/* ---- RDS data-link and application layer ----------------------------------------------------
 * The reference stops at printing syndrome matches (frame_thread, src/fm_radio.cpp:625-718).  This layer takes the
 * differentially decoded bits the chain emits (fmrx_outputs.rds_bits) and does what IEC 62106 / the course spec names
 * as the purpose of the RDS path: block synchronisation on the offset words A, B, C, C', D at 26-bit spacing (acquired on
 * two error-free blocks 26 bits apart in cyclic order, lost after 12 consecutive blocks that were not error-free),
 * checkword verification with burst-error correction (bursts <= 5 bits, shortened cyclic code (26,16), g(x) = 0x5B9;
 * attempted only while fewer than 3 consecutive blocks failed the clean check), group assembly, and decoding of PI, PTY, TP, the programme-service name (groups 0A/0B) and RadioText (2A/2B).
 * Host code (1187.5 bit/s per station is not GPU work); one opaque handle serves n_streams stations. */
typedef struct fmrx_rds_app fmrx_rds_app;
typedef struct {
    uint16_t blk[4];    /* information words of blocks A, B, C (or C'), D */
    uint8_t type;       /* group type 0..15 */
    uint8_t version_b;  /* 0 = version A, 1 = version B (block 3 carried offset C') */
    uint8_t corrected;  /* number of blocks of this group that needed error correction */
    uint8_t reserved;
    uint32_t bit_index; /* index, in the stream of fed bits, of the first bit of block A */
} fmrx_rds_group;
typedef struct {
    int32_t synced;           /* 1 while block synchronisation is held */
    int32_t pi;               /* programme identification, -1 until a block A was accepted */
    int32_t pty, tp;          /* programme type 0..31, traffic-programme flag; -1 until known */
    char ps[9];               /* programme-service name, 8 chars + NUL; '_' where no segment arrived yet */
    char rt[65];              /* RadioText, up to 64 chars + NUL, cut at the 0x0D terminator; '_' where unknown */
    uint8_t ps_complete, rt_ab_flag;
    uint32_t groups, blocks_ok, blocks_corrected, blocks_bad, sync_losses;
    uint64_t bits_fed;
} fmrx_rds_station;
int fmrx_rds_app_create(int n_streams, fmrx_rds_app **out);
void fmrx_rds_app_destroy(fmrx_rds_app *);
int fmrx_rds_app_reset(fmrx_rds_app *);
/* bits:[S][n_blocks][FMRX_MAX_BITS] and n_bits:[S][n_blocks] exactly as a process call returns them (host memory);
 * groups:[S][cap] receives the groups completed by this call (may be NULL), n_groups:[S] their count (may be NULL;
 * groups beyond `cap` are still decoded into the station record, only not returned) */
int fmrx_rds_app_feed(fmrx_rds_app *, const uint8_t *bits, const int32_t *n_bits, int n_blocks, fmrx_rds_group *groups, int cap, int32_t *n_groups);
int fmrx_rds_app_station(const fmrx_rds_app *, int stream, fmrx_rds_station *out);

/* ---- model-compatible operators ------------------------------------------------------------
 * The arithmetic of the reference's PYTHON models (the model directory), float64, so that fmMonoBlock.py / fmRDSblock.py can be
 * diffed tightly against the GPU; the C++ program is numerically a different receiver.  Batched over
 * n_streams independent streams ([stream][sample]); host buffers. */
/* scipy.signal.firwin(ntaps, cutoff, window='hann') as the models call it: one cutoff (fraction of Nyquist) with
 * pass_zero=1 = low-pass, two with pass_zero=0 = band-pass (model/fmMonoBlock.py:43-45,115,150,159) */
int fmrx_model_firwin(int ntaps, const double *cutoff, int n_cutoff, int pass_zero, double *h);
/* scipy.signal.lfilter(b, 1.0, x, zi=...) then the models' [::decim] slicing, on the input zero-stuffed by `up`
 * (model/fmRDSblock.py:188-199); y:[S][n*up/decim].  hist:[S][ntaps-1] = the last ntaps-1 input samples (zeros before
 * the first block), updated in place.  Products and sums are rounded one by one, oldest tap first (a transposed direct
 * form II); scipy's own FIR shortcut sums in numpy.convolve's order, so agreement is 1e-13 relative, not bit for bit. */
int fmrx_model_lfilter(double *y, const double *x, int n_streams, int n, const double *b, int ntaps, double *hist, int decim, int up);
/* fmSupportLib.fmDemodArctan (model/fmSupportLib.py:12-44): atan2 + numpy.unwrap against the carried phase; prev_phase:[S] in/out */
int fmrx_model_demod(double *demod, const double *I, const double *Q, int n_streams, int n, double *prev_phase);
/* fmPll.fmPll (model/fmPll.py:4-56): nco, nco_q:[S][n+1] (untrimmed like the model's; nco_q[0] is 0 where the model leaves
 * it uninitialised); state:[S][6] in the MODEL's order: integrator, phaseEst, feedbackI, feedbackQ, ncoOut[0], trigOffset */
int fmrx_model_pll(double *nco, double *nco_q, const double *x, int n_streams, int n, double freq, double Fs, double nco_scale, double phase_adjust,
                   double norm_bandwidth, double *state);

/* ---- measurement helpers (used by bench.py; not part of the receive path) ---------------------------------- */
/* runs an FP32 issue-rate microbenchmark on `device` and returns the best-of-`reps` rate in T lane-ops/s:
 * kind 0 = FFMA, 1 = FMUL+FADD pairs, 2 = packed FFMA2, 3 = packed FMUL2+FADD2; and the pipes the PLL step leans on:
 * 4 = DFMA (FP64), 5 = float<->double conversions (F2F pairs), 6 = integer ALU (SHF+LOP3 pairs) */
int fmrx_measure_fp32_peak(int device, int kind, int reps, double *tera_ops_per_s);
/* latency roofline of the PLL kernel: SM cycles per step of ONE loop's dependency chain (phase detector -> loop filter ->
 * oscillator, src/helper.cpp:32-45 as csrc/fmrx_pllmath.h computes it) run alone on one warp from registers */
int fmrx_measure_pll_chain(int device, double *cycles_per_step);

#ifdef __cplusplus
}
#endif
#endif /* FMRX_H */
