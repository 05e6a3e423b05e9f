// Internal declarations shared by the translation units of libfmrx.so: error plumbing and the kernel launchers.
// All launchers take DEVICE pointers, enqueue on `st` and return a cudaError_t-compatible int (0 = ok).
#ifndef FMRX_INTERNAL_H
#define FMRX_INTERNAL_H
#include <cstddef>
#include <cstdint>

#include "fmrx.h"

struct CUstream_st;
typedef CUstream_st *fmrx_stream_t;

namespace fmrx {

int fail(int status, const char *fmt, ...);
long long &launch_counter();  // process-wide count of kernels launched by this library

constexpr int kTaps = FMRX_MAX_TAPS;      // 151
constexpr int kHist = FMRX_MAX_TAPS - 1;  // 150 samples of live history

// how the FIR input sample at logical position p of block b is formed (p < 0 = history)
enum SrcKind {
    SRC_PLAIN = 0,    // x[p]; history one sample late (Q1)
    SRC_SQUARE = 1,   // x[p]^2; zi holds squares (pllCombine, src/helper.cpp:139,162-164)
    SRC_MIX_LATE = 2, // a[p]*b[p]; history one sample late (stereo mixer + LPF, src/fm_radio.cpp:269-274)
    SRC_MIX_HALF = 3, // 2*a[p]*b[p] in the block, a*b (not late, no x2) in the history (Q8, src/filter.cpp:387,399)
    SRC_PROD_HALF = 4 // the same filter fed with the ready-made product p = a*b (the PLL kernel writes it next to its NCO
                      // output): staged as p in the block and p/2 in the history, the sum doubled at the end -- scaling by two
                      // is exact at every step, so this is the MIX_HALF sum with one input signal instead of two
};

struct FirJob {
    const float *x;   // [S][ldx]: n_blocks*n samples per stream
    const float *x2;  // second factor for the MIX kinds (same layout), else null
    float *y;         // [S][ldy]: n_blocks*(n/decim)
    float *zi;        // [S][nzi]; the last 150 entries are live
    const float *h;   // HOST pointer to 151 taps (passed to the kernel by value)
    long long ldx, ldy;
    int nzi, n, n_blocks, n_streams, decim, kind, exact;
    int live_state_only;  // 1: the state update writes only the entries a later call can read (the last 150); the chain sets it --
                          // mode 1 sizes these states at 3623 (src/fm_radio.cpp:189-193) and nothing ever reads the rest
};
int launch_fir(const FirJob &j, fmrx_stream_t st);

struct FirMultiJob {  // 2 or 3 plain D = 1 filters of ONE input in one launch (same results as separate launch_fir calls, one staging)
    const float *x;       // [S][ldx]
    float *y[3];          // [S][ldy] each
    float *zi[3];         // [S][nzi] each; all carry the input's tail (they are updated together)
    const float *h[3];    // HOST taps
    long long ldx, ldy;
    int nzi, n, n_blocks, n_streams, nf, exact;  // n must be a multiple of 1024
};
int launch_fir_multi(const FirMultiJob &j, fmrx_stream_t st);

struct FirIqJob {  // two channels sharing taps (convolveWithDecimIQ)
    const float *xi, *xq;
    float *yi, *yq, *zii, *ziq;
    const float *h;
    long long ldx, ldy;
    int n, n_blocks, n_streams, decim, exact;
};
int launch_fir_iq(const FirIqJob &j, fmrx_stream_t st);

struct FrontendJob {  // rf_thread fused: u8 IQ -> FIR(151) /10 on I and Q -> discriminator
    const uint8_t *raw;  // [S][ld_raw] bytes: n_blocks * 2n per stream
    float *demod;        // [S][ld_out]
    float *yi, *yq;      // optional filtered I/Q, same layout as demod
    float *zii, *ziq;    // [S][150]
    const float *h;      // HOST taps
    long long ld_raw, ld_out;
    int n, n_blocks, n_streams;  // n complex samples per block; decim is fixed at 10
};
int launch_frontend(const FrontendJob &j, fmrx_stream_t st);

int launch_unpack(const uint8_t *raw, size_t n, float *out, fmrx_stream_t st);
int launch_demod(const float *I, const float *Q, float *out, int n_streams, int n_blocks, int n, fmrx_stream_t st);

struct ResampleJob {
    const float *x;  // [S][ldx]
    float *y;        // [S][ldy]
    float *zi;       // [S][nzi]
    const float *h;  // DEVICE taps [ntaps]
    const float *h_host;  // the same taps in HOST memory, or null: lets the tiled fast path (fmrx_resample.cu) pass them by value
    const float *hp;      // DEVICE, optional: the taps phase-major, hp[ph * 152 + c] = h[ph + up * c] (151 taps per phase): enables the phase-grouped kernel
    long long ldx, ldy;
    // n = samples per block in memory; n_ref = the length the reference's vector had (differs only for the RDS resampler,
    // whose input is the 15361-long mixer output, src/fm_radio.cpp:404-408); ny = outputs per block actually produced
    int n, n_ref, ny, n_blocks, n_streams, ntaps, nzi, decim, up, gain_up, exact;
    int live_state_only;  // 1: update only the state entries the history map (nzi-1-c)/up, c <= 150, can reach (Q6); set by the chain
};
int launch_resample(const ResampleJob &j, fmrx_stream_t st);
// register-tiled fast path for the RDS 19/80 geometry; returns -1 (nothing launched) when `j` is any other job
int launch_resample_tiled(const ResampleJob &j, fmrx_stream_t st);

struct PllParams {
    float freq, Fs, scale, phase_adj, bw;
};
// x,nco: [S][ld] with n_blocks*n samples; state [S][6].  One launch runs up to two independent PLL populations
// (stereo pilot + RDS carrier); pass xb == nullptr for a single one.
// mula/mulb (optional, same layout): a second signal whose product with the NCO OUTPUT is written to proda/prodb in the same
// pass (the stereo mixer src/fm_radio.cpp:269-272 and the RDS mixer src/filter.cpp:387 without its x2), so that the
// filters behind the PLLs read one signal instead of two.
int launch_pll_blocks(const float *xa, float *ncoa, PllParams pa, float *sta, const float *xb, float *ncob, PllParams pb, float *stb,
                      long long ld, int n_streams, int n, int n_blocks, fmrx_stream_t st, const float *mula = nullptr, float *proda = nullptr,
                      const float *mulb = nullptr, float *prodb = nullptr);

// mixed = a*b elementwise (mode-1 stereo path materialises it for the resampler, src/fm_radio.cpp:240-245)
int launch_multiply(const float *a, const float *b, float *y, long long ld, int n_total, int n_streams, fmrx_stream_t st);

struct CombineJob {  // L/R combine + quantise, src/fm_radio.cpp:277-299
    const float *mono, *stereo;  // [S][ld]; stereo may be null (binary profile after block 0, Q7)
    int16_t *audio;              // [S][2*ld] or null
    float *audio_f;              // [S][2*ld] or null
    long long ld;
    int n_total, n_streams, mult;
    int mono_delay;              // quality profile: mono taken this many samples late (<= 16) ...
    float *mono_tail;            // ... with the previous call's last samples here, [S][16]; null = no delay
};
int launch_combine(const CombineJob &j, fmrx_stream_t st);
// de-emphasis + quantiser on [S][n_blocks][2n] float audio (in place when write_float), state [S][4] = x[-1], y[-1] of L then R
int launch_deemphasis(float *audio_f, int16_t *audio, long long ld2, int n, int n_blocks, int n_streams, double b, double a1, int mult, float *state, int write_float,
                      fmrx_stream_t st);

// rrc [S][ld] with n_blocks*n; outputs as in fmrx_rds_decode; state [S][FMRX_RDS_STATE_WORDS]
int launch_rds_decode(const float *rrc, long long ld, int n_streams, int n_blocks, int n, uint8_t *bits, int32_t *n_bits,
                      fmrx_rds_event *events, int32_t *n_events, int32_t *state, fmrx_stream_t st);

// RDS back end at symbol rate (fmrx_rdsfast.cu): mixer LPF, 19/80 resampler and RRC as one composite polyphase filter
// evaluated only at the 152 samples per block the decoder reads; the reference's block-edge semantics enter through a second
// table over two windows of the previous block's products
struct RdsFastJob {
    const float *p;                 // [S][ld] mixer product NCO x RDS band (no x2), n_blocks * 15360 per station
    float *rrc;                     // [S][ldr] RRC buffer: written at 24k + off (and 0..23 in the very first block)
    float *edge;                    // carried state [S][edge_stride]: the 1094 edge products of the last block processed
    const float *W, *E;             // DEVICE: composite tables from rds_fast_tables
    const int32_t *off_state;       // DEVICE: sampling phase of station s at off_state[s * off_state_stride] (decoder state), valid when !first_block_is_zero
    int32_t *off_scratch;           // DEVICE [S]: receives the phase derived from block 0 when first_block_is_zero
    long long ld, ldr;
    int off_state_stride, n_streams, n_blocks, first_block_is_zero, edge_stride;
};
int rds_fast_tables(const float *h1, const float *h2, const float *hr, float **dW, float **dE);
int launch_rds_fast(const RdsFastJob &j, fmrx_stream_t st);

int measure_fp32_peak(int device, int kind, int reps, double *tera);
int measure_pll_chain(double *cycles_per_step);
// shape of a batch handle, for the host-side layers that sit on top of the C-ABI (fmrx_ring.cpp)
void batch_shape(const fmrx_batch *b, int *n_streams, int *max_blocks, int *audio_per_block, int *audio_on, int *rds_on);

// SM partition (green contexts) for the device-resident pipeline, fmrx_partition.cu.  partition_create returns nullptr
// when the driver cannot split the device; the caller then falls back to priority streams on the whole device.
struct SmPartition;
SmPartition *partition_create(int device, int want_small, int prio_small, fmrx_stream_t *s_small, int n_big, const int *prio_big, fmrx_stream_t *s_big);
void partition_destroy(SmPartition *p);
void partition_sizes(const SmPartition *p, int *small, int *big);

}  // namespace fmrx
#endif
