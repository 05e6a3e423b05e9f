// C-ABI of libfmrx.so (include/fmrx.h): error plumbing, the function-level operators (host buffers in/out around the
// same kernel launchers the chain uses) and the batched receive chain with its CUDA-stream pipeline.
//
// The chain reproduces the four thread bodies of /root/reference/src/fm_radio.cpp:
//   rf_thread :31-147, mono_stero_thread :150-318, rds_thread :321-441, frame_thread :444-729
// as a fixed sequence of kernel launches per (chunk of streams x blocks).  The reference's producer/consumer threads
// and bounded queues (:86-138, :212-223, :376-387, :414-423) become: a copy-in stream, two compute streams and a
// copy-out stream, chained by events, working on chunks of the stream dimension so that H2D of chunk c+1, the kernels of
// chunk c and D2H of chunk c-1 overlap.  There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <atomic>
#include <vector>

#include "fmrx_internal.h"

namespace fmrx {

static thread_local char g_err[512] = "";

int fail(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

long long &launch_counter() {
    static long long c = 0;
    return c;
}

static int cuda_fail(cudaError_t e, const char *what) {
    return fail(e == cudaErrorMemoryAllocation ? FMRX_ERR_ALLOC : FMRX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(call)                                           \
    do {                                                   \
        cudaError_t e_ = (call);                           \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)
#define LAUNCH(call)                                                         \
    do {                                                                     \
        int e_ = (call);                                                     \
        if (e_ != 0) return cuda_fail(static_cast<cudaError_t>(e_), #call); \
    } while (0)

// RAII device buffer for the function-level entries
template <class T>
struct Dev {
    T *p = nullptr;
    size_t n = 0;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) { n = count; return cudaMalloc(&p, (count ? count : 1) * sizeof(T)); }
    cudaError_t up(const T *h, size_t count) { cudaError_t e = alloc(count); return e ? e : cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T *h, size_t count) const { return cudaMemcpy(h, p, count * sizeof(T), cudaMemcpyDeviceToHost); }
};

static int check_taps(const float *h, int ntaps) {
    if (!h || ntaps != kTaps) return fail(FMRX_ERR_ARG, "this operator is specialised for %d taps (got %d)", kTaps, ntaps);
    return FMRX_OK;
}

}  // namespace fmrx

using namespace fmrx;

extern "C" {

const char *fmrx_last_error(void) { return g_err; }
int fmrx_version(void) { return FMRX_VERSION; }
int fmrx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------------------------------------------------------
// function-level operators
// ------------------------------------------------------------------------------------------------------------------
int fmrx_unpack_iq(const uint8_t *raw, size_t n, float *out) {
    if (!raw || !out) return fail(FMRX_ERR_ARG, "fmrx_unpack_iq: null pointer");
    if (n == 0) return FMRX_OK;
    Dev<uint8_t> d_in; Dev<float> d_out;
    CU(d_in.up(raw, n)); CU(d_out.alloc(n));
    LAUNCH(launch_unpack(d_in.p, n, d_out.p, nullptr));
    CU(d_out.down(out, n));
    return FMRX_OK;
}

int fmrx_fir_decim(float *y, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps, float *zi, int nzi,
                   int decim, int exact) {
    if (!y || !x || !zi || n_streams <= 0 || n_blocks <= 0 || n < kTaps + 1 || nzi < kHist) return fail(FMRX_ERR_ARG, "fmrx_fir_decim: bad argument");
    if (decim != 1 && decim != 5 && decim != 10) return fail(FMRX_ERR_ARG, "fmrx_fir_decim: decim must be 1, 5 or 10");
    if (int e = check_taps(h, ntaps)) return e;
    const size_t nx = (size_t)n_streams * n_blocks * n, ny = (size_t)n_streams * n_blocks * (n / decim);
    Dev<float> dx, dy, dz;
    CU(dx.up(x, nx)); CU(dy.alloc(ny)); CU(dz.up(zi, (size_t)n_streams * nzi));
    FirJob j{};
    j.x = dx.p; j.y = dy.p; j.zi = dz.p; j.h = h; j.ldx = (long long)n_blocks * n; j.ldy = (long long)n_blocks * (n / decim);
    j.nzi = nzi; j.n = n; j.n_blocks = n_blocks; j.n_streams = n_streams; j.decim = decim; j.kind = SRC_PLAIN; j.exact = exact;
    LAUNCH(launch_fir(j, nullptr));
    CU(dy.down(y, ny)); CU(dz.down(zi, (size_t)n_streams * nzi));
    return FMRX_OK;
}

int fmrx_fir_decim_iq(float *yi, float *yq, const float *xi, const float *xq, int n_streams, int n_blocks, int n, const float *h,
                      int ntaps, float *zii, float *ziq, int decim, int exact) {
    if (!yi || !yq || !xi || !xq || !zii || !ziq || n_streams <= 0 || n_blocks <= 0 || n < kTaps + 1) return fail(FMRX_ERR_ARG, "fmrx_fir_decim_iq: bad argument");
    if (decim != 10) return fail(FMRX_ERR_ARG, "fmrx_fir_decim_iq: decim must be 10 (src/fm_radio.cpp:42)");
    if (int e = check_taps(h, ntaps)) return e;
    const size_t nx = (size_t)n_streams * n_blocks * n, ny = (size_t)n_streams * n_blocks * (n / decim), nz = (size_t)n_streams * kHist;
    Dev<float> dxi, dxq, dyi, dyq, dzi, dzq;
    CU(dxi.up(xi, nx)); CU(dxq.up(xq, nx)); CU(dyi.alloc(ny)); CU(dyq.alloc(ny)); CU(dzi.up(zii, nz)); CU(dzq.up(ziq, nz));
    FirIqJob j{};
    j.xi = dxi.p; j.xq = dxq.p; j.yi = dyi.p; j.yq = dyq.p; j.zii = dzi.p; j.ziq = dzq.p; j.h = h;
    j.ldx = (long long)n_blocks * n; j.ldy = (long long)n_blocks * (n / decim); j.n = n; j.n_blocks = n_blocks; j.n_streams = n_streams;
    j.decim = decim; j.exact = exact;
    LAUNCH(launch_fir_iq(j, nullptr));
    CU(dyi.down(yi, ny)); CU(dyq.down(yq, ny)); CU(dzi.down(zii, nz)); CU(dzq.down(ziq, nz));
    return FMRX_OK;
}

int fmrx_resample(float *y, int ny_limit, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps, float *zi,
                  int nzi, int decim, int up, int gain_up, int exact) {
    if (!y || !x || !h || !zi || n_streams <= 0 || n_blocks <= 0 || n <= 0 || ntaps <= 0 || nzi <= 0 || decim <= 0 || up <= 0)
        return fail(FMRX_ERR_ARG, "fmrx_resample: bad argument");
    if (n < nzi + 1) return fail(FMRX_ERR_ARG, "fmrx_resample: block (%d) must be longer than the state (%d): the reference's state update reads x[n-nzi-1]", n, nzi);
    const long long ny_full = ((long long)n * up) / decim;
    const int ny = (ny_limit > 0 && ny_limit < ny_full) ? ny_limit : (int)ny_full;
    const size_t nx = (size_t)n_streams * n_blocks * n, nyt = (size_t)n_streams * n_blocks * ny;
    Dev<float> dx, dy, dz, dh;
    CU(dx.up(x, nx)); CU(dy.alloc(nyt)); CU(dz.up(zi, (size_t)n_streams * nzi)); CU(dh.up(h, ntaps));
    ResampleJob j{};
    j.x = dx.p; j.y = dy.p; j.zi = dz.p; j.h = dh.p; j.h_host = h; j.ldx = (long long)n_blocks * n; j.ldy = (long long)n_blocks * ny;
    j.n = n; j.n_ref = n; j.ny = ny; j.n_blocks = n_blocks; j.n_streams = n_streams; j.ntaps = ntaps; j.nzi = nzi; j.decim = decim; j.up = up;
    j.gain_up = gain_up; j.exact = exact;
    LAUNCH(launch_resample(j, nullptr));
    CU(dy.down(y, nyt)); CU(dz.down(zi, (size_t)n_streams * nzi));
    return FMRX_OK;
}

int fmrx_fir_mixer(float *y, const float *nco, const float *sig, int n_streams, int n_blocks, int n, const float *h, int ntaps, float *zi) {
    if (!y || !nco || !sig || !zi || n_streams <= 0 || n_blocks <= 0 || n < kTaps + 1) return fail(FMRX_ERR_ARG, "fmrx_fir_mixer: bad argument");
    if (int e = check_taps(h, ntaps)) return e;
    const size_t nx = (size_t)n_streams * n_blocks * n, nz = (size_t)n_streams * kHist;
    Dev<float> da, db, dy, dz;
    CU(da.up(nco, nx)); CU(db.up(sig, nx)); CU(dy.alloc(nx)); CU(dz.up(zi, nz));
    FirJob j{};
    j.x = da.p; j.x2 = db.p; j.y = dy.p; j.zi = dz.p; j.h = h; j.ldx = j.ldy = (long long)n_blocks * n;
    j.nzi = kHist; j.n = n; j.n_blocks = n_blocks; j.n_streams = n_streams; j.decim = 1; j.kind = SRC_MIX_HALF; j.exact = 1;  // ((x*x1)*h)*2 == ((x*x1)*2)*h: scaling by two is exact
    LAUNCH(launch_fir(j, nullptr));
    CU(dy.down(y, nx)); CU(dz.down(zi, nz));
    return FMRX_OK;
}

int fmrx_demod(const float *I, const float *Q, int n_streams, int n_blocks, int n, float *out) {
    if (!I || !Q || !out || n_streams <= 0 || n_blocks <= 0 || n <= 0) return fail(FMRX_ERR_ARG, "fmrx_demod: bad argument");
    const size_t nx = (size_t)n_streams * n_blocks * n;
    Dev<float> di, dq, dout;
    CU(di.up(I, nx)); CU(dq.up(Q, nx)); CU(dout.alloc(nx));
    LAUNCH(launch_demod(di.p, dq.p, dout.p, n_streams, n_blocks, n, nullptr));
    CU(dout.down(out, nx));
    return FMRX_OK;
}

int fmrx_pll(float *nco, const float *x, int n_streams, int n_blocks, int n, float freq, float Fs, float scale, float phase_adj,
             float bw, float *state) {
    if (!nco || !x || !state || n_streams <= 0 || n_blocks <= 0 || n <= 0) return fail(FMRX_ERR_ARG, "fmrx_pll: bad argument");
    const size_t nx = (size_t)n_streams * n_blocks * n;
    Dev<float> dx, dn, ds;
    CU(dx.up(x, nx)); CU(dn.alloc(nx)); CU(ds.up(state, (size_t)n_streams * 6));
    PllParams p{freq, Fs, scale, phase_adj, bw};
    LAUNCH(launch_pll_blocks(dx.p, dn.p, p, ds.p, nullptr, nullptr, p, nullptr, (long long)n_blocks * n, n_streams, n, n_blocks, nullptr));
    CU(dn.down(nco, nx)); CU(ds.down(state, (size_t)n_streams * 6));
    return FMRX_OK;
}

int fmrx_pll_combine(float *y, float *nco, const float *x, int n_streams, int n_blocks, int n, const float *h, int ntaps, float *zi,
                     float freq, float Fs, float scale, float phase_adj, float bw, float *state) {
    if (!y || !nco || !x || !zi || !state || n_streams <= 0 || n_blocks <= 0 || n < kTaps + 1) return fail(FMRX_ERR_ARG, "fmrx_pll_combine: bad argument");
    if (int e = check_taps(h, ntaps)) return e;
    const size_t nx = (size_t)n_streams * n_blocks * n, nz = (size_t)n_streams * kHist;
    Dev<float> dx, dy, dn, dz, ds;
    CU(dx.up(x, nx)); CU(dy.alloc(nx)); CU(dn.alloc(nx)); CU(dz.up(zi, nz)); CU(ds.up(state, (size_t)n_streams * 6));
    FirJob j{};
    j.x = dx.p; j.y = dy.p; j.zi = dz.p; j.h = h; j.ldx = j.ldy = (long long)n_blocks * n;
    j.nzi = kHist; j.n = n; j.n_blocks = n_blocks; j.n_streams = n_streams; j.decim = 1; j.kind = SRC_SQUARE; j.exact = 1;  // double products into a float sum, src/helper.cpp:139
    LAUNCH(launch_fir(j, nullptr));
    PllParams p{freq, Fs, scale, phase_adj, bw};
    LAUNCH(launch_pll_blocks(dy.p, dn.p, p, ds.p, nullptr, nullptr, p, nullptr, (long long)n_blocks * n, n_streams, n, n_blocks, nullptr));
    CU(dy.down(y, nx)); CU(dn.down(nco, nx)); CU(dz.down(zi, nz)); CU(ds.down(state, (size_t)n_streams * 6));
    return FMRX_OK;
}

int fmrx_frontend(float *demod, float *yi, float *yq, const uint8_t *raw, int n_streams, int n_blocks, int n, const float *h, int ntaps,
                  float *zii, float *ziq, int decim) {
    if (!demod || !raw || !zii || !ziq || n_streams <= 0 || n_blocks <= 0 || n < kTaps + 1 || (yi == nullptr) != (yq == nullptr))
        return fail(FMRX_ERR_ARG, "fmrx_frontend: bad argument");
    if (decim != 10) return fail(FMRX_ERR_ARG, "fmrx_frontend: decim must be 10 (src/fm_radio.cpp:42)");
    if (int e = check_taps(h, ntaps)) return e;
    const size_t nraw = (size_t)n_streams * n_blocks * n * 2, ny = (size_t)n_streams * n_blocks * (n / 10), nz = (size_t)n_streams * kHist;
    Dev<uint8_t> draw; Dev<float> dd, dyi, dyq, dzi, dzq;
    CU(draw.up(raw, nraw)); CU(dd.alloc(ny)); CU(dzi.up(zii, nz)); CU(dzq.up(ziq, nz));
    if (yi) { CU(dyi.alloc(ny)); CU(dyq.alloc(ny)); }
    FrontendJob j{};
    j.raw = draw.p; j.demod = dd.p; j.yi = yi ? dyi.p : nullptr; j.yq = yi ? dyq.p : nullptr; j.zii = dzi.p; j.ziq = dzq.p; j.h = h;
    j.ld_raw = (long long)n_blocks * n * 2; j.ld_out = (long long)n_blocks * (n / 10); j.n = n; j.n_blocks = n_blocks; j.n_streams = n_streams;
    LAUNCH(launch_frontend(j, nullptr));
    CU(dd.down(demod, ny)); CU(dzi.down(zii, nz)); CU(dzq.down(ziq, nz));
    if (yi) { CU(dyi.down(yi, ny)); CU(dyq.down(yq, ny)); }
    return FMRX_OK;
}

int fmrx_deemphasis(float *audio_f, int16_t *audio, int n_streams, int n_blocks, int n, float tau_us, float Fs, int mult, float *state) {
    if (!audio_f || !state || n_streams <= 0 || n_blocks <= 0 || n <= 0 || n > 24000) return fail(FMRX_ERR_ARG, "fmrx_deemphasis: bad argument");
    double bb, a1;
    if (int e = fmrx_deemphasis_coeffs(tau_us, Fs, &bb, &a1)) return e;
    const size_t nx = (size_t)n_streams * n_blocks * 2 * n;
    Dev<float> df, ds; Dev<int16_t> da;
    CU(df.up(audio_f, nx)); CU(ds.up(state, (size_t)n_streams * 4));
    if (audio) CU(da.alloc(nx));
    LAUNCH(launch_deemphasis(df.p, audio ? da.p : nullptr, (long long)n_blocks * 2 * n, n, n_blocks, n_streams, bb, a1, mult, ds.p, 1, nullptr));
    CU(df.down(audio_f, nx)); CU(ds.down(state, (size_t)n_streams * 4));
    if (audio) CU(da.down(audio, nx));
    return FMRX_OK;
}

int fmrx_rds_decode(const float *rrc, int n_streams, int n_blocks, int n, uint8_t *bits, int32_t *n_bits, fmrx_rds_event *events,
                    int32_t *n_events, int32_t *state) {
    if (!rrc || !state || n_streams <= 0 || n_blocks <= 0 || n < 24 * 8 || n / 24 / 2 + 1 > FMRX_MAX_BITS) return fail(FMRX_ERR_ARG, "fmrx_rds_decode: bad argument");
    const size_t nx = (size_t)n_streams * n_blocks * n, nb = (size_t)n_streams * n_blocks;
    Dev<float> dx; Dev<uint8_t> dbits; Dev<int32_t> dnb, dne, dst; Dev<fmrx_rds_event> dev;
    CU(dx.up(rrc, nx)); CU(dbits.alloc(nb * FMRX_MAX_BITS)); CU(dnb.alloc(nb)); CU(dne.alloc(nb)); CU(dev.alloc(nb * FMRX_MAX_EVENTS));
    CU(dst.up(state, (size_t)n_streams * FMRX_RDS_STATE_WORDS));
    CU(cudaMemset(dbits.p, 0, nb * FMRX_MAX_BITS));
    LAUNCH(launch_rds_decode(dx.p, (long long)n_blocks * n, n_streams, n_blocks, n, dbits.p, dnb.p, dev.p, dne.p, dst.p, nullptr));
    if (bits) CU(dbits.down(bits, nb * FMRX_MAX_BITS));
    if (n_bits) CU(dnb.down(n_bits, nb));
    if (events) CU(dev.down(events, nb * FMRX_MAX_EVENTS));
    if (n_events) CU(dne.down(n_events, nb));
    CU(dst.down(state, (size_t)n_streams * FMRX_RDS_STATE_WORDS));
    return FMRX_OK;
}

int fmrx_rds_state_offset(const int32_t *state) { return state ? state[1] : -1; }

int fmrx_rds_format_block(int block_id, int initial_offset, const fmrx_rds_event *ev, int n_ev, char *buf, int cap) {
    // the literal lines of src/fm_radio.cpp:516, :619-620, :652-701 (including the reference's spelling)
    int w = 0;
    auto emit = [&](const char *fmt, auto... a) {
        int r = snprintf(buf && w < cap ? buf + w : nullptr, buf && w < cap ? (size_t)(cap - w) : 0, fmt, a...);
        if (r > 0) w += r;
    };
    if (block_id == 0) emit("initial offset for clock recovery = %d\n", initial_offset);
    emit(" \n****************Prcoessing Block: %d****************\n", block_id);
    for (int i = 0; i < n_ev; ++i) {
        if (ev[i].kind == FMRX_EV_RESYNC) emit("~~~~~Re-Sync~~~~~\n");
        else emit("%sSyndrome %c at position %u\n", ev[i].kind == FMRX_EV_FALSE ? "False positive " : "", "ABCD"[ev[i].letter & 3], ev[i].position);
    }
    return w;
}

int fmrx_measure_pll_chain(int device, double *cycles_per_step) {
    if (!cycles_per_step) return fail(FMRX_ERR_ARG, "null pointer");
    CU(cudaSetDevice(device));
    LAUNCH(measure_pll_chain(cycles_per_step));
    return FMRX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// the batched chain
// ------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int NIQ = FMRX_BLOCK_BYTES / 2;  // complex samples per block
constexpr int NIF = FMRX_IF_PER_BLOCK;
constexpr int NRDS = FMRX_RDS_PER_BLOCK;
constexpr double kPi = 3.14159265358979323846;
constexpr int kMaxChunks = 8;
constexpr int kSets = 3;       // rotating buffer sets of the device-resident pipeline
constexpr int kTickets = 8;    // completion events kept by the asynchronous host path

}  // namespace

struct fmrx_batch {
    fmrx_config cfg{};
    int S = 0, NB = 0, n_audio = 0, nzi_a = 0, nzi_b = 0, audio_taps = 0, up = 1, decim_a = 5, mult = 1;
    bool audio_on = false, rds_on = false, exact = true;
    bool strict = false;    // FMRX_NUMERICS_STRICT: the squared-input filter in front of the 114 kHz PLL with the reference's double products, staged RDS back end exact
    bool rds_fast = false;  // RDS back end at symbol rate (fmrx_rdsfast.cu) instead of stage by stage
    float *d_W = nullptr, *d_E = nullptr;
    int32_t *d_off = nullptr;
    long long block_id = 0;  // blocks consumed per stream so far
    int last_blocks = 0;
    long long launches = 0;
    // host taps
    float h_rf[kTaps], h_pilot[kTaps], h_sbpf[kTaps], h_rbpf[kTaps], h_sq[kTaps], h_lpf3k[kTaps], h_rrc[kTaps];
    std::vector<float> h_mono, h_stereo, h_anti;
    float rds_phase = 0.f;
    // quality profile (FMRX_QUALITY_*): de-emphasis coefficients (b == 0: off), mono delay in audio samples (0: off) and their states
    double de_b = 0.0, de_a1 = 0.0;
    int mono_delay = 0;
    float *mono_tail = nullptr, *deemph_st = nullptr;
    // device
    float *d_h_mono = nullptr, *d_h_stereo = nullptr, *d_h_anti = nullptr, *d_hp_mono = nullptr, *d_hp_stereo = nullptr;
    char *d_state = nullptr;  // one blob: every carried state
    size_t state_bytes = 0;
    float *zi_i, *zi_q, *zi_mono, *zi_pilot, *zi_sbpf, *zi_stereo, *pll_st, *zi_rbpf, *zi_sq, *zi_lpf, *zi_rrc, *zi_anti, *rds_pll_st;
    int32_t *dec_st;
    uint8_t *d_iq = nullptr;
    float *demod = nullptr, *mono = nullptr, *pilot = nullptr, *nco = nullptr, *sbpf = nullptr, *mixed = nullptr, *stereo = nullptr;
    float *rbpf = nullptr, *rsq = nullptr, *rnco = nullptr, *rmixed = nullptr, *rlpf = nullptr, *rres = nullptr, *rrrc = nullptr;
    int16_t *audio = nullptr;
    float *audio_f = nullptr;
    uint8_t *bits = nullptr;
    int32_t *nbits = nullptr, *nev = nullptr;
    fmrx_rds_event *ev = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr, s_cmp[2] = {nullptr, nullptr};
    cudaEvent_t e_in[kMaxChunks]{}, e_done[kMaxChunks]{}, e_out[kMaxChunks]{};
    // device-resident pipeline: filters before the PLLs, the PLLs, and everything after them run on three streams, so
    // that step k's (latency-bound, few-warp) PLL kernel overlaps step k+1's front end and step k-1's back end.  The
    // signals crossing a phase boundary are kept in kSets rotating buffer sets (`set`): with two, phase A of call k+2
    // waits for phase C of call k, which waits for call k's PLLs -- every other step the PLL partition idled for the length
    // of phase A (fmrx_batch_timeline); with three the only steady-state limit is the slowest phase.
    // When the batch is large enough to fill the device, the PLL stream lives in a green context with SMs of its own
    // and the two filter streams in a second one with the rest (fmrx_partition.cu); s_ser is a whole-device stream for
    // the serialised per-stage profiling pass.
    cudaStream_t s_a = nullptr, s_p = nullptr, s_c = nullptr, s_ser = nullptr;
    fmrx::SmPartition *part = nullptr;
    cudaEvent_t ev_a[kSets]{}, ev_p[kSets]{}, ev_c[kSets]{};
    bool ev_c_valid[kSets] = {};
    bool was_serial = false;
    // asynchronous host path (fmrx_batch_submit / fmrx_batch_wait): two device staging buffers for the IQ bytes, so that
    // the H2D of step k+1 runs under the kernels of step k; tickets are events on the copy-out stream
    uint8_t *d_iq2 = nullptr;
    cudaEvent_t e_h2d[2]{}, e_iqfree[2]{}, e_d2h = nullptr, ev_ticket[kTickets]{};
    bool iqfree_valid[2] = {false, false}, d2h_valid = false;
    bool want_audio_f = false;  // set per call: combine_kernel writes the float audio only when the caller takes it
    std::atomic<long long> submits{0};  // read by fmrx_batch_wait, which a consumer thread may call while a producer thread submits (fmrx_ring)
    long long calls = 0;
    int last_set = 0;
    size_t set_if = 0, set_au = 0;  // elements per set of an IF-rate / audio-rate signal
    std::vector<void *> allocs;
    // optional per-stage device timing (fmrx_batch_profile): event pairs around every stage of every enqueued chain
    bool profiling = false;
    bool profile_pipelined = false;  // fmrx_batch_profile(.., 2): keep the three-stream pipeline while timing (timeline mode)
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    struct Mark { int stage; cudaEvent_t t0, t1; };
    std::vector<Mark> marks;
    double stage_ms[FMRX_STAGE_COUNT] = {0};
    long long stage_launches[FMRX_STAGE_COUNT] = {0};

    ~fmrx_batch() {
        cudaSetDevice(cfg.device);
        for (void *p : allocs) cudaFree(p);
        for (auto e : prof_pool) cudaEventDestroy(e);
        for (int i = 0; i < kMaxChunks; ++i) { if (e_in[i]) cudaEventDestroy(e_in[i]); if (e_done[i]) cudaEventDestroy(e_done[i]); if (e_out[i]) cudaEventDestroy(e_out[i]); }
        for (int i = 0; i < kSets; ++i) { if (ev_a[i]) cudaEventDestroy(ev_a[i]); if (ev_p[i]) cudaEventDestroy(ev_p[i]); if (ev_c[i]) cudaEventDestroy(ev_c[i]); }
        for (auto e : e_h2d) if (e) cudaEventDestroy(e);
        for (auto e : e_iqfree) if (e) cudaEventDestroy(e);
        for (auto e : ev_ticket) if (e) cudaEventDestroy(e);
        if (e_d2h) cudaEventDestroy(e_d2h);
        for (auto st : {s_a, s_p, s_c, s_ser}) if (st) cudaStreamDestroy(st);
        fmrx::partition_destroy(part);
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        for (auto s : s_cmp) if (s) cudaStreamDestroy(s);
    }
    template <class T>
    cudaError_t dalloc(T *&p, size_t count) {
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (!e) allocs.push_back(p);
        return e;
    }
};

namespace {

int init_state(fmrx_batch *b) {
    // reference initial conditions: all filter histories 0 (src/fm_radio.cpp:53-55,189-193,360-364), both PLLs
    // {0,0,1,0,0,1} (:165-171,:343-349), decoder zeroed (:449-484)
    CU(cudaMemsetAsync(b->d_state, 0, b->state_bytes, b->s_cmp[0]));
    std::vector<float> pll((size_t)b->S * 6);
    for (int s = 0; s < b->S; ++s) { float *p = &pll[(size_t)s * 6]; p[0] = 0; p[1] = 0; p[2] = 1; p[3] = 0; p[4] = 0; p[5] = 1; }
    CU(cudaMemcpyAsync(b->pll_st, pll.data(), pll.size() * 4, cudaMemcpyHostToDevice, b->s_cmp[0]));
    CU(cudaMemcpyAsync(b->rds_pll_st, pll.data(), pll.size() * 4, cudaMemcpyHostToDevice, b->s_cmp[0]));
    CU(cudaStreamSynchronize(b->s_cmp[0]));
    b->block_id = 0;
    b->calls = 0;
    for (bool &v : b->ev_c_valid) v = false;
    b->iqfree_valid[0] = b->iqfree_valid[1] = b->d2h_valid = false;
    return FMRX_OK;
}

struct StageScope {
    fmrx_batch *b; cudaStream_t st; int stage; cudaEvent_t t1 = nullptr;
    StageScope(fmrx_batch *b_, cudaStream_t st_, int stage_) : b(b_), st(st_), stage(stage_) {
        if (!b->profiling) return;
        auto get = [&]() { if (b->prof_used == b->prof_pool.size()) { cudaEvent_t e; cudaEventCreate(&e); b->prof_pool.push_back(e); } return b->prof_pool[b->prof_used++]; };
        cudaEvent_t t0 = get(); t1 = get();
        cudaEventRecord(t0, st);
        b->marks.push_back({stage, t0, t1});
    }
    ~StageScope() { if (t1) cudaEventRecord(t1, st); }
};
#define STAGE(id) StageScope stage_scope_##id(b, st, id)

// enqueue the whole chain for streams [s0, s0+ns) x nblk blocks: phase A (everything before the PLLs) on stA, the PLLs
// on stP, phase C (everything after) on stC.  With three distinct streams the phases are chained by ev_a/ev_p of `set`.
int enqueue_chain(fmrx_batch *b, const uint8_t *iq, long long ld_iq, int s0, int ns, int nblk, const fmrx_outputs &o, int set,
                  cudaStream_t stA, cudaStream_t stP, cudaStream_t stC) {
    const long long before = launch_counter();
    const long long ldif = (long long)b->NB * NIF, lda = (long long)b->NB * b->n_audio, ldr = (long long)b->NB * NRDS;
    auto IF = [&](float *p) { return p + (long long)s0 * ldif; };                                // single-buffered
    auto IF2 = [&](float *p) { return p + (long long)set * (long long)b->set_if + (long long)s0 * ldif; };  // crosses a phase boundary
    auto AU = [&](float *p) { return p + (long long)s0 * lda; };
    auto AU2 = [&](float *p) { return p + (long long)set * (long long)b->set_au + (long long)s0 * lda; };
    auto RD = [&](float *p) { return p + (long long)s0 * ldr; };
    const bool piped = stA != stP;
    cudaStream_t st = stA;
    const int ex = b->exact ? 1 : 0;
    // ---- rf_thread: unpack + deinterleave + LPF/10 + discriminator (src/fm_radio.cpp:66-84)
    FrontendJob f{};
    f.raw = iq + (long long)s0 * ld_iq; f.demod = IF(b->demod); f.zii = b->zi_i + (long long)s0 * kHist; f.ziq = b->zi_q + (long long)s0 * kHist; f.h = b->h_rf;
    f.ld_raw = ld_iq; f.ld_out = ldif; f.n = NIQ; f.n_blocks = nblk; f.n_streams = ns;
    { STAGE(FMRX_STAGE_FRONTEND); LAUNCH(launch_frontend(f, st)); }

    auto fir = [&](const float *x, const float *x2, float *y, float *zi, int nzi, const float *h, long long ldx, long long ldy, int n, int decim, int kind, int exact, int blocks) {
        FirJob j{};
        j.x = x; j.x2 = x2; j.y = y; j.zi = zi; j.h = h; j.ldx = ldx; j.ldy = ldy; j.nzi = nzi; j.n = n; j.n_blocks = blocks; j.n_streams = ns;
        j.decim = decim; j.kind = kind; j.exact = exact; j.live_state_only = 1;
        return launch_fir(j, st);
    };
    const bool stereo_live = b->cfg.profile == FMRX_PROFILE_INTENT || b->block_id == 0;  // Q7
    const int st_blocks = b->cfg.profile == FMRX_PROFILE_INTENT ? nblk : 1;

    // ---- mono_stero_thread, filters (:226-236 / :255-265)
    if (b->audio_on) {
        STAGE(FMRX_STAGE_MONO);
        if (b->cfg.mode != 0) {
            ResampleJob r{};
            r.x = IF(b->demod); r.y = AU2(b->mono); r.zi = b->zi_mono + (long long)s0 * b->nzi_a; r.h = b->d_h_mono; r.ldx = ldif; r.ldy = lda;
            r.n = NIF; r.n_ref = NIF; r.ny = b->n_audio; r.n_blocks = nblk; r.n_streams = ns; r.ntaps = b->audio_taps; r.nzi = b->nzi_a;
            r.decim = b->decim_a; r.up = b->up; r.gain_up = 0; r.exact = ex; r.hp = b->d_hp_mono; r.live_state_only = 1;
            LAUNCH(launch_resample(r, st));
        } else {
            LAUNCH(fir(IF(b->demod), nullptr, AU2(b->mono), b->zi_mono + (long long)s0 * b->nzi_a, b->nzi_a, b->h_mono.data(), ldif, lda, NIF, 5, SRC_PLAIN, ex, nblk));
        }
    }
    // the three band-pass filters of the discriminator output -- pilot 18.5-19.5 kHz (:232/:261), stereo band 22-54 kHz (:236/:265),
    // RDS band 54-60 kHz (:395) -- can share ONE launch when they run over the same blocks with the same rounding (intent profile:
    // every block; their states are then updated together and hold the same input tail): the tile is staged once instead of three
    // times (fir151_multi_kernel).  Measured (4096 stations): the fused-multiply-add pair 0.566 ms against 2 x 0.292 separately --
    // adopted; the three exact filters 1.818 ms against 3 x 0.575 -- NOT adopted: with a run-time filter index the 151 taps reach the
    // uniform registers as 158 scalar LDCU instead of the 50 wide ones of the single-filter kernel, +6 % instructions in a kernel at
    // 90 % of its issue roof, more than the two stagings it saves.  FMRX_BPF_FUSED=1 forces the fused form for the exact filters too.
    static const bool force_fused = [] { const char *v = getenv("FMRX_BPF_FUSED"); return v && atoi(v) == 1; }();
    const bool same_blocks = b->audio_on && stereo_live && st_blocks == nblk && NIF % 1024 == 0;
    const bool fuse3 = force_fused && same_blocks && b->rds_on && b->exact && b->nzi_b == kHist;
    const bool fuse2_audio = force_fused && same_blocks && b->exact && !b->rds_on;      // pilot + stereo band, both exact
    const bool fuse2_fma = same_blocks && !b->exact && b->rds_on && b->nzi_b == kHist;  // stereo band + RDS band, both FMA
    if (fuse3 || fuse2_audio || fuse2_fma) {
        FirMultiJob m{};
        m.x = IF(b->demod); m.ldx = ldif; m.ldy = ldif; m.n = NIF; m.n_blocks = nblk; m.n_streams = ns; m.nzi = b->nzi_b; m.exact = fuse2_fma ? 0 : 1;
        int nf = 0;
        auto add = [&](float *y, float *zi, const float *h) { m.y[nf] = y; m.zi[nf] = zi; m.h[nf] = h; ++nf; };
        if (!fuse2_fma) add(IF2(b->pilot), b->zi_pilot + (long long)s0 * b->nzi_b, b->h_pilot);
        add(IF2(b->sbpf), b->zi_sbpf + (long long)s0 * b->nzi_b, b->h_sbpf);
        if (fuse3 || fuse2_fma) add(IF2(b->rbpf), b->zi_rbpf + (long long)s0 * kHist, b->h_rbpf);
        m.nf = nf;
        if (fuse2_fma) { STAGE(FMRX_STAGE_PILOT_BPF); LAUNCH(fir(IF(b->demod), nullptr, IF2(b->pilot), b->zi_pilot + (long long)s0 * b->nzi_b, b->nzi_b, b->h_pilot, ldif, ldif, NIF, 1, SRC_PLAIN, 1, st_blocks)); }
        STAGE(FMRX_STAGE_BPF_FUSED);
        LAUNCH(launch_fir_multi(m, st));
    } else {
        if (b->audio_on && stereo_live) {
            { STAGE(FMRX_STAGE_PILOT_BPF); LAUNCH(fir(IF(b->demod), nullptr, IF2(b->pilot), b->zi_pilot + (long long)s0 * b->nzi_b, b->nzi_b, b->h_pilot, ldif, ldif, NIF, 1, SRC_PLAIN, 1, st_blocks)); }
            STAGE(FMRX_STAGE_STEREO_BPF);
            LAUNCH(fir(IF(b->demod), nullptr, IF2(b->sbpf), b->zi_sbpf + (long long)s0 * b->nzi_b, b->nzi_b, b->h_sbpf, ldif, ldif, NIF, 1, SRC_PLAIN, ex, st_blocks));
        }
        // REFERENCE / STRICT: the 54-60 kHz band-pass with the reference's two roundings per tap (src/filter.cpp:126-154)
        if (b->rds_on) { STAGE(FMRX_STAGE_RDS_BPF); LAUNCH(fir(IF(b->demod), nullptr, IF2(b->rbpf), b->zi_rbpf + (long long)s0 * kHist, kHist, b->h_rbpf, ldif, ldif, NIF, 1, SRC_PLAIN, ex, nblk)); }
    }
    // ---- rds_thread, pllCombine's filter (:400).  STRICT: with its double products (src/helper.cpp:139), so that the 114 kHz loop sees
    // the reference's input bit for bit
    if (b->rds_on) {
        STAGE(FMRX_STAGE_RDS_SQ_BPF);
        LAUNCH(fir(IF2(b->rbpf), nullptr, IF2(b->rsq), b->zi_sq + (long long)s0 * kHist, kHist, b->h_sq, ldif, ldif, NIF, 1, SRC_SQUARE, b->strict ? 1 : 0, nblk));
    }
    // ---- both PLLs, one lane per (stream, loop) (:233/:262 and :400)
    if (piped) { CU(cudaEventRecord(b->ev_a[set], stA)); CU(cudaStreamWaitEvent(stP, b->ev_a[set], 0)); }
    st = stP;
    {
        const PllParams pa{19e3f, 240e3f, 2.0f, 0.0f, 0.01f};
        const PllParams pr{114000.0f, 240000.0f, 0.5f, b->rds_phase, 0.001f};
        const bool a_on = b->audio_on && stereo_live;
        STAGE(FMRX_STAGE_PLL);
        // each loop also writes its NCO output times the signal it will be mixed with (stereo band / RDS band), so the
        // filters of phase C read one signal instead of two
        if (a_on && b->rds_on && st_blocks == nblk) {
            LAUNCH(launch_pll_blocks(IF2(b->pilot), IF2(b->nco), pa, b->pll_st + (long long)s0 * 6, IF2(b->rsq), IF2(b->rnco), pr, b->rds_pll_st + (long long)s0 * 6, ldif, ns, NIF, nblk, st,
                                     IF2(b->sbpf), IF2(b->mixed), IF2(b->rbpf), IF2(b->rmixed)));
        } else {
            if (a_on) LAUNCH(launch_pll_blocks(IF2(b->pilot), IF2(b->nco), pa, b->pll_st + (long long)s0 * 6, nullptr, nullptr, pa, nullptr, ldif, ns, NIF, st_blocks, st,
                                               IF2(b->sbpf), IF2(b->mixed)));
            if (b->rds_on) LAUNCH(launch_pll_blocks(IF2(b->rsq), IF2(b->rnco), pr, b->rds_pll_st + (long long)s0 * 6, nullptr, nullptr, pr, nullptr, ldif, ns, NIF, nblk, st,
                                                    IF2(b->rbpf), IF2(b->rmixed)));
        }
    }
    if (piped) { CU(cudaEventRecord(b->ev_p[set], stP)); CU(cudaStreamWaitEvent(stC, b->ev_p[set], 0)); }
    st = stC;
    // ---- stereo mix + LPF, combine, quantise (:240-252 / :269-299)
    if (b->audio_on) {
        const float *stereo = nullptr;
        if (stereo_live) {
            STAGE(FMRX_STAGE_STEREO_LPF);
            if (st_blocks < nblk) CU(cudaMemset2DAsync(AU(b->stereo), lda * 4, 0, (size_t)nblk * b->n_audio * 4, ns, st));
            if (b->cfg.mode != 0) {
                ResampleJob r{};  // convolveWithDecimMode1(stereo_filt, mixed, stereo_coeff, stereo_initial, 5, 24), :245 (Q14); mode 2: the decimation it meant
                r.x = IF2(b->mixed); r.y = AU(b->stereo); r.zi = b->zi_stereo + (long long)s0 * b->nzi_a; r.h = b->d_h_stereo; r.ldx = ldif; r.ldy = lda;
                r.n = NIF; r.n_ref = NIF; r.ny = b->n_audio; r.n_blocks = st_blocks; r.n_streams = ns; r.ntaps = b->audio_taps; r.nzi = b->nzi_a;
                r.decim = b->cfg.mode == 1 ? 5 : b->decim_a; r.up = b->up; r.gain_up = 0; r.exact = ex; r.hp = b->d_hp_stereo; r.live_state_only = 1;
                LAUNCH(launch_resample(r, st));
            } else {
                LAUNCH(fir(IF2(b->mixed), nullptr, AU(b->stereo), b->zi_stereo + (long long)s0 * b->nzi_a, b->nzi_a, b->h_stereo.data(), ldif, lda, NIF, 5, SRC_PLAIN, ex, st_blocks));
            }
            stereo = AU(b->stereo);
        }
        STAGE(FMRX_STAGE_COMBINE);
        CombineJob c{};
        c.mono = AU2(b->mono); c.stereo = stereo; c.audio = b->audio + (long long)s0 * lda * 2; c.audio_f = b->want_audio_f ? b->audio_f + (long long)s0 * lda * 2 : nullptr;  // the float copy is 2/3 of this kernel's writes: only when asked for
        c.ld = lda; c.n_total = nblk * b->n_audio; c.n_streams = ns; c.mult = b->mult;
        c.mono_delay = b->mono_delay; c.mono_tail = b->mono_delay ? b->mono_tail + (long long)s0 * 16 : nullptr;
        const bool deemph = b->de_b != 0.0;
        if (deemph) { c.audio = nullptr; c.audio_f = b->audio_f + (long long)s0 * lda * 2; }  // the quantiser moves behind the de-emphasis
        LAUNCH(launch_combine(c, st));
        if (deemph)
            LAUNCH(launch_deemphasis(b->audio_f + (long long)s0 * lda * 2, b->audio + (long long)s0 * lda * 2, lda * 2, b->n_audio, nblk, ns, b->de_b, b->de_a1, b->mult,
                                     b->deemph_st + (long long)s0 * 4, b->want_audio_f ? 1 : 0, st));
    }
    // ---- rds_thread after the PLL (:404-411) and frame_thread
    if (b->rds_on) {
        if (b->rds_fast) {
            RdsFastJob q{};
            // the symbol-rate path keeps ONE state, the edge products of the last block, in the segment the staged path uses for the
            // resampler history (2868 floats per station; the mixer and RRC history segments stay unused)
            q.p = IF2(b->rmixed); q.rrc = RD(b->rrrc); q.edge = b->zi_anti + (long long)s0 * (kTaps * 19 - 1); q.edge_stride = kTaps * 19 - 1; q.W = b->d_W; q.E = b->d_E;
            q.off_state = b->dec_st + (long long)s0 * FMRX_RDS_STATE_WORDS + 1; q.off_state_stride = FMRX_RDS_STATE_WORDS; q.off_scratch = b->d_off + s0;
            q.ld = ldif; q.ldr = ldr; q.n_streams = ns; q.n_blocks = nblk; q.first_block_is_zero = b->block_id == 0;
            STAGE(FMRX_STAGE_RDS_SYMBOLS);
            LAUNCH(launch_rds_fast(q, st));
        } else {
            { STAGE(FMRX_STAGE_RDS_MIX_LPF); LAUNCH(fir(IF2(b->rmixed), nullptr, IF(b->rlpf), b->zi_lpf + (long long)s0 * kHist, kHist, b->h_lpf3k, ldif, ldif, NIF, 1, SRC_PROD_HALF, b->strict ? 1 : 0, nblk)); }
            ResampleJob r{};
            r.x = IF(b->rlpf); r.y = RD(b->rres); r.zi = b->zi_anti + (long long)s0 * (kTaps * 19 - 1); r.h = b->d_h_anti; r.h_host = b->h_anti.data(); r.ldx = ldif; r.ldy = ldr;
            r.n = NIF; r.n_ref = NIF + 1; r.ny = NRDS; r.n_blocks = nblk; r.n_streams = ns; r.ntaps = kTaps * 19; r.nzi = kTaps * 19 - 1;
            r.decim = 80; r.up = 19; r.gain_up = 1; r.exact = b->strict ? 1 : 0;
            { STAGE(FMRX_STAGE_RDS_RESAMPLE); LAUNCH(launch_resample(r, st)); }
            { STAGE(FMRX_STAGE_RDS_RRC); LAUNCH(fir(RD(b->rres), nullptr, RD(b->rrrc), b->zi_rrc + (long long)s0 * kHist, kHist, b->h_rrc, ldr, ldr, NRDS, 1, SRC_PLAIN, b->strict ? 1 : 0, nblk)); }
        }
        STAGE(FMRX_STAGE_RDS_DECODE);
        LAUNCH(launch_rds_decode(RD(b->rrrc), ldr, ns, nblk, NRDS, b->bits + (long long)s0 * nblk * FMRX_MAX_BITS, b->nbits + (long long)s0 * nblk,
                                 b->ev + (long long)s0 * nblk * FMRX_MAX_EVENTS, b->nev + (long long)s0 * nblk, b->dec_st + (long long)s0 * FMRX_RDS_STATE_WORDS, st));
    }
    (void)o;
    b->launches += launch_counter() - before;
    b->last_set = set;
    return FMRX_OK;
}

// copy one chunk's results to the caller (device->device or device->host depending on `kind`)
int copy_outputs(fmrx_batch *b, int s0, int ns, int nblk, const fmrx_outputs &o, cudaMemcpyKind kind, cudaStream_t st) {
    const size_t lda = (size_t)b->NB * b->n_audio, row = (size_t)nblk * b->n_audio;
    if (b->audio_on && o.audio)
        CU(cudaMemcpy2DAsync(o.audio + (size_t)s0 * row * 2, row * 2 * sizeof(int16_t), b->audio + (size_t)s0 * lda * 2, lda * 2 * sizeof(int16_t), row * 2 * sizeof(int16_t), ns, kind, st));
    if (b->audio_on && o.audio_f)
        CU(cudaMemcpy2DAsync(o.audio_f + (size_t)s0 * row * 2, row * 2 * sizeof(float), b->audio_f + (size_t)s0 * lda * 2, lda * 2 * sizeof(float), row * 2 * sizeof(float), ns, kind, st));
    if (b->rds_on) {
        const size_t q = (size_t)s0 * nblk, m = (size_t)ns * nblk;
        if (o.rds_bits) CU(cudaMemcpyAsync(o.rds_bits + q * FMRX_MAX_BITS, b->bits + q * FMRX_MAX_BITS, m * FMRX_MAX_BITS, kind, st));
        if (o.rds_n_bits) CU(cudaMemcpyAsync(o.rds_n_bits + q, b->nbits + q, m * sizeof(int32_t), kind, st));
        if (o.rds_events) CU(cudaMemcpyAsync(o.rds_events + q * FMRX_MAX_EVENTS, b->ev + q * FMRX_MAX_EVENTS, m * FMRX_MAX_EVENTS * sizeof(fmrx_rds_event), kind, st));
        if (o.rds_n_events) CU(cudaMemcpyAsync(o.rds_n_events + q, b->nev + q, m * sizeof(int32_t), kind, st));
    }
    return FMRX_OK;
}

}  // namespace

namespace fmrx {
void batch_shape(const fmrx_batch *b, int *n_streams, int *max_blocks, int *audio_per_block, int *audio_on, int *rds_on) {
    *n_streams = b->S; *max_blocks = b->NB; *audio_per_block = b->n_audio; *audio_on = b->audio_on ? 1 : 0; *rds_on = b->rds_on ? 1 : 0;
}
}  // namespace fmrx

extern "C" {

int fmrx_batch_create(const fmrx_config *cfg, fmrx_batch **out) {
    if (!cfg || !out) return fail(FMRX_ERR_ARG, "fmrx_batch_create: null pointer");
    *out = nullptr;
    if (cfg->mode < 0 || cfg->mode > 2 || cfg->numerics < 0 || cfg->numerics > 2 || cfg->n_streams <= 0 || cfg->max_blocks <= 0 || cfg->n_streams > 65535 || cfg->max_blocks > 65535)
        return fail(FMRX_ERR_ARG, "fmrx_batch_create: mode must be 0, 1 or 2, 1 <= n_streams,max_blocks <= 65535");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(FMRX_ERR_CUDA, "no CUDA device: libfmrx has no CPU fallback"); }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FMRX_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));
    fmrx_batch *b = new (std::nothrow) fmrx_batch();
    if (!b) return fail(FMRX_ERR_ALLOC, "out of host memory");
    struct Guard { fmrx_batch *p; ~Guard() { delete p; } } guard{b};
    b->cfg = *cfg;
    b->S = cfg->n_streams; b->NB = cfg->max_blocks;
    const int paths = cfg->paths ? cfg->paths : (FMRX_PATH_AUDIO | FMRX_PATH_RDS);
    b->audio_on = (paths & FMRX_PATH_AUDIO) != 0;
    b->rds_on = (paths & FMRX_PATH_RDS) != 0 && cfg->mode != 1;  // src/fm_radio.cpp:324,446 (mode 2 has mode 0's 240 kHz IF)
    b->rds_fast = b->rds_on && !(paths & FMRX_PATH_RDS_STAGES);
    b->exact = cfg->numerics != FMRX_NUMERICS_FMA;
    b->strict = cfg->numerics == FMRX_NUMERICS_STRICT;
    // ---- constants of the thread bodies
    const float rf_Fs = cfg->mode == 1 ? 2500000.0f : 2400000.0f;  // :36-37
    float audio_Fs = 240000.0f;                                    // :153
    b->audio_taps = kTaps; b->up = 1; b->decim_a = 5; b->mult = 1;
    if (cfg->mode == 1) { audio_Fs = 6000000.0f; b->decim_a = 125; b->up = 24; b->audio_taps = kTaps * 24; b->mult = 24; }  // :174-180,229
    // mode 2 (not in the reference): the mode-1 thread body with (U, D) = (147, 800) on mode 0's 240 kHz IF -> 44.1 kHz, 2822 per block
    if (cfg->mode == 2) { audio_Fs = 240000.0f * 147.0f; b->decim_a = 800; b->up = 147; b->audio_taps = kTaps * 147; b->mult = 147; }
    b->nzi_a = b->audio_taps - 1;                                  // :189-193
    if (b->nzi_a > NIF - 1) b->nzi_a = NIF - 1;                    // mode 2 only: the reference's update rule zi[i] = x[N - Z - 1 + i] needs Z <= N - 1
    b->nzi_b = cfg->mode == 2 ? kHist : b->nzi_a;                  // the two 151-tap band-pass filters: :189-193 size them from audio_taps too; mode 2 keeps the 150 live entries
    b->n_audio = (int)(((long long)NIF * b->up) / b->decim_a);
    b->h_mono.resize(b->audio_taps); b->h_stereo.resize(b->audio_taps); b->h_anti.resize(kTaps * 19);
    fmrx_design_lpf(rf_Fs, 100000.0f, kTaps, b->h_rf);                                  // :40-42,75
    fmrx_design_lpf(audio_Fs, 16000.0f, (unsigned short)b->audio_taps, b->h_mono.data());   // :200
    const float bpf_Fs = cfg->mode == 2 ? 240000.0f : audio_Fs;  // :201-202 use audio_Fs, in mode 1 the upsampled 6 MHz (replicated); mode 2: the IF rate, like mode 0
    fmrx_design_bpf(18.5e3f, 19.5e3f, bpf_Fs, kTaps, b->h_pilot);                       // :201
    fmrx_design_bpf(22e3f, 54e3f, bpf_Fs, kTaps, b->h_sbpf);                            // :202
    fmrx_design_lpf(audio_Fs, 16000.0f, (unsigned short)b->audio_taps, b->h_stereo.data()); // :203
    fmrx_design_bpf(54000.0f, 60000.0f, 240000.0f, kTaps, b->h_rbpf);                   // :366
    fmrx_design_bpf(113500.0f, 114500.0f, 240000.0f, kTaps, b->h_sq);                   // :367
    fmrx_design_lpf(240000.0f, 3000.0f, kTaps, b->h_lpf3k);                             // :368
    fmrx_design_lpf(240000.0f * 19.0f, (float)(57000 / 2), kTaps * 19, b->h_anti.data()); // :369
    fmrx_design_rrc(57000.0f, kTaps, b->h_rrc);                                         // :370
    const float phase_adj = (float)(kPi / 3.3 - kPi / 1.5);                             // :342
    b->rds_phase = (float)((double)phase_adj - kPi / 1.4);                              // :400
    // ---- quality profile: never the default (the output is no longer the reference's)
    if (cfg->quality & ~(FMRX_QUALITY_DEEMPH_75 | FMRX_QUALITY_DEEMPH_50 | FMRX_QUALITY_UNITY_BPF | FMRX_QUALITY_AUTO_RDS_PHASE))
        return fail(FMRX_ERR_ARG, "fmrx_batch_create: unknown quality flag in %d", cfg->quality);
    if ((cfg->quality & FMRX_QUALITY_DEEMPH_75) && (cfg->quality & FMRX_QUALITY_DEEMPH_50)) return fail(FMRX_ERR_ARG, "fmrx_batch_create: one de-emphasis time constant, not two");
    if (cfg->quality & FMRX_QUALITY_UNITY_BPF) {
        fmrx_design_bpf_unity(18.5e3f, 19.5e3f, bpf_Fs, kTaps, b->h_pilot);
        fmrx_design_bpf_unity(22e3f, 54e3f, bpf_Fs, kTaps, b->h_sbpf);
        fmrx_design_bpf_unity(54000.0f, 60000.0f, 240000.0f, kTaps, b->h_rbpf);
        fmrx_design_bpf_unity(113500.0f, 114500.0f, 240000.0f, kTaps, b->h_sq);
        for (float &v : b->h_stereo) v *= 2.0f;  // the x2 of the stereo mixer, folded into the (linear) low-pass behind it: exact
        // the L-R branch spends 75 IF samples more in filters than L+R: in audio samples 75 * U / D (15 in mode 0, rounded otherwise)
        b->mono_delay = (int)((75LL * b->up + b->decim_a / 2) / b->decim_a);
    }
    if (cfg->quality & FMRX_QUALITY_AUTO_RDS_PHASE) fmrx_rds_auto_phase(b->h_sq, kTaps, 240000.0f, 114000.0f, &b->rds_phase);
    if (cfg->quality & (FMRX_QUALITY_DEEMPH_75 | FMRX_QUALITY_DEEMPH_50)) {
        const float audio_rate = (float)((double)(cfg->mode == 1 ? 250000.0 : 240000.0) * b->up / b->decim_a);
        fmrx_deemphasis_coeffs((cfg->quality & FMRX_QUALITY_DEEMPH_75) ? 75.0f : 50.0f, audio_rate, &b->de_b, &b->de_a1);
    }

    const size_t S = b->S, NB = b->NB;
    b->set_if = S * NB * NIF; b->set_au = S * NB * b->n_audio;
    // ---- carried state: one blob
    struct Seg { void **p; size_t bytes; };
    std::vector<Seg> segs = {
        {(void **)&b->zi_i, S * kHist * 4}, {(void **)&b->zi_q, S * kHist * 4},
        {(void **)&b->zi_mono, S * b->nzi_a * 4}, {(void **)&b->zi_pilot, S * b->nzi_b * 4}, {(void **)&b->zi_sbpf, S * b->nzi_b * 4}, {(void **)&b->zi_stereo, S * b->nzi_a * 4},
        {(void **)&b->pll_st, S * 6 * 4}, {(void **)&b->zi_rbpf, S * kHist * 4}, {(void **)&b->zi_sq, S * kHist * 4}, {(void **)&b->zi_lpf, S * kHist * 4},
        {(void **)&b->zi_rrc, S * kHist * 4}, {(void **)&b->zi_anti, S * (kTaps * 19 - 1) * 4}, {(void **)&b->rds_pll_st, S * 6 * 4},
        {(void **)&b->dec_st, S * FMRX_RDS_STATE_WORDS * 4}, {(void **)&b->mono_tail, S * 16 * 4}, {(void **)&b->deemph_st, S * 4 * 4}};
    size_t total = 0;
    for (auto &sg : segs) total += (sg.bytes + 255) & ~(size_t)255;
    b->state_bytes = total;
    CU(b->dalloc(b->d_state, total));
    size_t off = 0;
    for (auto &sg : segs) { *sg.p = b->d_state + off; off += (sg.bytes + 255) & ~(size_t)255; }
    // ---- taps of the long (resampler) filters
    CU(b->dalloc(b->d_h_mono, b->audio_taps)); CU(b->dalloc(b->d_h_stereo, b->audio_taps)); CU(b->dalloc(b->d_h_anti, kTaps * 19));
    CU(cudaMemcpy(b->d_h_mono, b->h_mono.data(), b->audio_taps * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b->d_h_stereo, b->h_stereo.data(), b->audio_taps * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(b->d_h_anti, b->h_anti.data(), kTaps * 19 * 4, cudaMemcpyHostToDevice));
    if (b->up > 1) {  // phase-major copies for the phase-grouped resampler kernel
        std::vector<float> hp((size_t)b->up * (kTaps + 1), 0.0f);
        for (int which = 0; which < 2; ++which) {
            const std::vector<float> &h = which ? b->h_stereo : b->h_mono;
            for (int ph = 0; ph < b->up; ++ph)
                for (int cc = 0; cc < kTaps; ++cc) hp[(size_t)ph * (kTaps + 1) + cc] = h[ph + (size_t)b->up * cc];
            float *&dst = which ? b->d_hp_stereo : b->d_hp_mono;
            CU(b->dalloc(dst, hp.size()));
            CU(cudaMemcpy(dst, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
        }
    }
    // ---- signals
    CU(b->dalloc(b->d_iq, S * NB * FMRX_BLOCK_BYTES));
    CU(b->dalloc(b->demod, S * NB * NIF));
    if (b->audio_on) {
        CU(b->dalloc(b->mono, kSets * S * NB * b->n_audio)); CU(b->dalloc(b->pilot, kSets * S * NB * NIF)); CU(b->dalloc(b->nco, kSets * S * NB * NIF));
        CU(b->dalloc(b->sbpf, kSets * S * NB * NIF)); CU(b->dalloc(b->stereo, S * NB * b->n_audio));
        CU(b->dalloc(b->mixed, kSets * S * NB * NIF));  // stereo-band x NCO, written by the PLL kernel
        CU(b->dalloc(b->audio, S * NB * b->n_audio * 2)); CU(b->dalloc(b->audio_f, S * NB * b->n_audio * 2));
    }
    if (b->rds_on) {
        CU(b->dalloc(b->rbpf, kSets * S * NB * NIF)); CU(b->dalloc(b->rsq, kSets * S * NB * NIF)); CU(b->dalloc(b->rnco, kSets * S * NB * NIF)); CU(b->dalloc(b->rmixed, kSets * S * NB * NIF)); CU(b->dalloc(b->rlpf, S * NB * NIF));
        CU(b->dalloc(b->rres, S * NB * NRDS)); CU(b->dalloc(b->rrrc, S * NB * NRDS));
        CU(b->dalloc(b->bits, S * NB * FMRX_MAX_BITS)); CU(b->dalloc(b->nbits, S * NB)); CU(b->dalloc(b->nev, S * NB)); CU(b->dalloc(b->ev, S * NB * FMRX_MAX_EVENTS));
        if (b->rds_fast) {
            CU((cudaError_t)fmrx::rds_fast_tables(b->h_lpf3k, b->h_anti.data(), b->h_rrc, &b->d_W, &b->d_E));
            b->allocs.push_back(b->d_W); b->allocs.push_back(b->d_E);
            CU(b->dalloc(b->d_off, S));
            CU(cudaMemset(b->rrrc, 0, S * NB * NRDS * sizeof(float)));  // only the decoder's positions are ever written
        }
        CU(cudaMemset(b->bits, 0, S * NB * FMRX_MAX_BITS));
        CU(cudaMemset(b->ev, 0, S * NB * FMRX_MAX_EVENTS * sizeof(fmrx_rds_event)));
    }
    CU(cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking)); CU(cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking));
    for (auto &s : b->s_cmp) CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    {
        // the PLL kernel is a few hundred single-warp CTAs that run for the whole step: give its stream the highest
        // priority so those CTAs are placed as soon as any slot frees up instead of queueing behind the FIR grids
        int least = 0, greatest = 0;
        CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        if (const char *e = getenv("FMRX_PIPE_PRIO")) { if (e[0] == '0') greatest = least; }  // experiment switch: equal priorities
        const int mid = greatest < least ? greatest + 1 : least;
        // FMRX_PLL_SMS: SMs set aside for the PLL stream (0 = no partition).  Default: the split that minimises
        // max(PLL phase, filter phases) under the measured cost model: the PLL kernel takes 3.2 / 4.4 / 6.1 / 8.8 ms
        // per 64 ms block at 1 / 2 / 3 / 4 warps per scheduler (one warp = 32 loops), the filters of 4096 stations take
        // f ms on the whole device and scale with the SMs they are left with.  4096 stations, stereo + RDS: 32 / 116.
        int pll_sms = 0;
        // binary profile: the stereo loop runs in block 0 only (Q7) -- without the RDS loop there is nothing to give a partition to
        const bool pilot_loop = b->audio_on && cfg->profile == FMRX_PROFILE_INTENT;
        if (b->S >= 1024 && (pilot_loop || b->rds_on)) {
            const int loops = b->S * ((pilot_loop ? 1 : 0) + (b->rds_on ? 1 : 0)), warps = (loops + 31) / 32;
            const double f_audio = cfg->mode == 0 ? 1.52 : cfg->mode == 1 ? 2.35 : 2.97;
            const double f_ms = (1.43 + (b->audio_on ? f_audio : 0.0) + (b->rds_on ? 1.09 : 0.0)) * b->S / 4096.0;
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
            double best = 1e30;
            for (int p = 8; p <= 64 && p < sms; p += 8) {
                const int w = (warps + 4 * p - 1) / (4 * p);
                const double t_pll = w <= 1 ? 3.2 : w == 2 ? (p >= 32 ? 4.4 : 5.0) : w == 3 ? 6.1 : w == 4 ? 8.8 : 2.2 * w;  // two warps per scheduler on a 16-SM partition (mode 1's single loop): 4.9-5.0 measured
                const double t = std::max(t_pll, f_ms * sms / (sms - p));
                if (t < best * 0.98) { best = t; pll_sms = p; }   // ties go to the smaller partition
            }
        }
        if (const char *e = getenv("FMRX_PLL_SMS")) pll_sms = atoi(e);
        if (pll_sms > 0) {
            const int prio_big[2] = {least, mid};
            cudaStream_t big[2] = {nullptr, nullptr};
            b->part = fmrx::partition_create(cfg->device, pll_sms, greatest, &b->s_p, 2, prio_big, big);
            if (b->part) { b->s_a = big[0]; b->s_c = big[1]; }
        }
        if (!b->part) {
            CU(cudaStreamCreateWithPriority(&b->s_p, cudaStreamNonBlocking, greatest));
            CU(cudaStreamCreateWithPriority(&b->s_c, cudaStreamNonBlocking, mid));
            CU(cudaStreamCreateWithPriority(&b->s_a, cudaStreamNonBlocking, least));
        }
        CU(cudaStreamCreateWithFlags(&b->s_ser, cudaStreamNonBlocking));
    }
    for (int i = 0; i < kSets; ++i) {
        CU(cudaEventCreateWithFlags(&b->ev_a[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b->ev_p[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b->ev_c[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < kMaxChunks; ++i) {
        CU(cudaEventCreateWithFlags(&b->e_in[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b->e_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b->e_out[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; ++i) { CU(cudaEventCreateWithFlags(&b->e_h2d[i], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&b->e_iqfree[i], cudaEventDisableTiming)); }
    CU(cudaEventCreateWithFlags(&b->e_d2h, cudaEventDisableTiming));
    for (auto &e : b->ev_ticket) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (int e = init_state(b)) return e;
    guard.p = nullptr;
    *out = b;
    return FMRX_OK;
}

void fmrx_batch_destroy(fmrx_batch *b) { delete b; }
int fmrx_batch_audio_per_block(const fmrx_batch *b) { return b ? b->n_audio : 0; }
long long fmrx_batch_launch_count(const fmrx_batch *b) { return b ? b->launches : 0; }
int fmrx_batch_partition(const fmrx_batch *b, int *pll_sms, int *filter_sms) {
    if (!b || !pll_sms || !filter_sms) return fail(FMRX_ERR_ARG, "null pointer");
    fmrx::partition_sizes(b->part, pll_sms, filter_sms);
    return FMRX_OK;
}
void *fmrx_batch_cuda_stream(fmrx_batch *b) { return b ? (void *)b->s_c : nullptr; }
void *fmrx_batch_cuda_stream_phase(fmrx_batch *b, int phase) {
    if (!b) return nullptr;
    return (void *)(phase == 0 ? b->s_a : phase == 1 ? b->s_p : b->s_c);
}

int fmrx_batch_reset(fmrx_batch *b) {
    if (!b) return fail(FMRX_ERR_ARG, "null handle");
    CU(cudaSetDevice(b->cfg.device));
    if (int e = fmrx_batch_sync(b)) return e;
    return init_state(b);
}

int fmrx_batch_sync(fmrx_batch *b) {
    if (!b) return fail(FMRX_ERR_ARG, "null handle");
    CU(cudaSetDevice(b->cfg.device));
    CU(cudaStreamSynchronize(b->s_in)); CU(cudaStreamSynchronize(b->s_cmp[0])); CU(cudaStreamSynchronize(b->s_cmp[1])); CU(cudaStreamSynchronize(b->s_out));
    CU(cudaStreamSynchronize(b->s_a)); CU(cudaStreamSynchronize(b->s_p)); CU(cudaStreamSynchronize(b->s_c)); CU(cudaStreamSynchronize(b->s_ser));
    return FMRX_OK;
}

int fmrx_batch_process_device(fmrx_batch *b, const uint8_t *iq_device, int n_blocks, const fmrx_outputs *out_device) {
    if (!b || !iq_device) return fail(FMRX_ERR_ARG, "fmrx_batch_process_device: null pointer");
    if (n_blocks <= 0 || n_blocks > b->NB) return fail(FMRX_ERR_ARG, "n_blocks %d outside 1..%d", n_blocks, b->NB);
    CU(cudaSetDevice(b->cfg.device));
    fmrx_outputs none{};
    const fmrx_outputs &o = out_device ? *out_device : none;
    const int set = (int)(b->calls % kSets);
    // phase A of this call overwrites the buffer set phase C of the call kSets back was reading
    if (b->ev_c_valid[set] && !(b->profiling && !b->profile_pipelined)) CU(cudaStreamWaitEvent(b->s_a, b->ev_c[set], 0));
    // while per-stage profiling is on, the three phases are serialised on one stream so that every stage is timed alone
    const bool serial = b->profiling && !b->profile_pipelined;
    if (serial != b->was_serial) {  // switching between the pipeline streams and the whole-device stream: drain first
        if (int e = fmrx_batch_sync(b)) return e;
        b->was_serial = serial;
    }
    cudaStream_t sa = serial ? b->s_ser : b->s_a, sp = serial ? b->s_ser : b->s_p, sc = serial ? b->s_ser : b->s_c;
    b->want_audio_f = o.audio_f != nullptr;
    if (b->d2h_valid) CU(cudaStreamWaitEvent(sc, b->e_d2h, 0));  // a copy-out of fmrx_batch_submit may still be reading the (single) result buffers
    if (int e = enqueue_chain(b, iq_device, (long long)n_blocks * FMRX_BLOCK_BYTES, 0, b->S, n_blocks, o, set, sa, sp, sc)) return e;
    if (int e = copy_outputs(b, 0, b->S, n_blocks, o, cudaMemcpyDeviceToDevice, sc)) return e;
    CU(cudaEventRecord(b->ev_c[set], sc));
    b->ev_c_valid[set] = true;
    b->calls += 1;
    b->block_id += n_blocks;
    b->last_blocks = n_blocks;
    return FMRX_OK;
}

int fmrx_batch_process(fmrx_batch *b, const uint8_t *iq, int n_blocks, const fmrx_outputs *out) {
    if (!b || !iq) return fail(FMRX_ERR_ARG, "fmrx_batch_process: null pointer");
    if (n_blocks <= 0 || n_blocks > b->NB) return fail(FMRX_ERR_ARG, "n_blocks %d outside 1..%d", n_blocks, b->NB);
    CU(cudaSetDevice(b->cfg.device));
    fmrx_outputs none{};
    const fmrx_outputs &o = out ? *out : none;
    if (b->calls) { if (int e = fmrx_batch_sync(b)) return e; }  // drain the device-resident pipeline first
    // chunk the stream dimension so copies and kernels overlap; small batches go through in one piece
    const int chunks = b->S >= 64 ? (b->S >= 512 ? kMaxChunks : 4) : 1;
    const long long row = (long long)n_blocks * FMRX_BLOCK_BYTES;
    for (int c = 0; c < chunks; ++c) {
        const int s0 = (int)((long long)b->S * c / chunks), s1 = (int)((long long)b->S * (c + 1) / chunks), ns = s1 - s0;
        if (ns <= 0) continue;
        cudaStream_t cs = b->s_cmp[c & 1];
        CU(cudaMemcpyAsync(b->d_iq + (size_t)s0 * row, iq + (size_t)s0 * row, (size_t)ns * row, cudaMemcpyHostToDevice, b->s_in));
        CU(cudaEventRecord(b->e_in[c], b->s_in));
        CU(cudaStreamWaitEvent(cs, b->e_in[c], 0));
        b->want_audio_f = o.audio_f != nullptr;
        if (int e = enqueue_chain(b, b->d_iq, row, s0, ns, n_blocks, o, 0, cs, cs, cs)) return e;
        CU(cudaEventRecord(b->e_done[c], cs));
        CU(cudaStreamWaitEvent(b->s_out, b->e_done[c], 0));
        if (int e = copy_outputs(b, s0, ns, n_blocks, o, cudaMemcpyDeviceToHost, b->s_out)) return e;
    }
    CU(cudaStreamSynchronize(b->s_out)); CU(cudaStreamSynchronize(b->s_cmp[0])); CU(cudaStreamSynchronize(b->s_cmp[1]));
    b->block_id += n_blocks;
    b->last_blocks = n_blocks;
    return FMRX_OK;
}

int fmrx_batch_submit(fmrx_batch *b, const uint8_t *iq, int n_blocks, const fmrx_outputs *out, long long *ticket) {
    if (!b || !iq || !ticket) return fail(FMRX_ERR_ARG, "fmrx_batch_submit: null pointer");
    if (n_blocks <= 0 || n_blocks > b->NB) return fail(FMRX_ERR_ARG, "n_blocks %d outside 1..%d", n_blocks, b->NB);
    if (b->profiling && !b->profile_pipelined) return fail(FMRX_ERR_STATE, "fmrx_batch_submit is not available while serialised stage profiling is on");
    CU(cudaSetDevice(b->cfg.device));
    fmrx_outputs none{};
    const fmrx_outputs &o = out ? *out : none;
    if (b->was_serial) { if (int e = fmrx_batch_sync(b)) return e; b->was_serial = false; }
    if (!b->d_iq2) CU(b->dalloc(b->d_iq2, (size_t)b->S * b->NB * FMRX_BLOCK_BYTES));
    const long long nsub = b->submits.load();
    const int slot = (int)(nsub & 1);
    uint8_t *dst = slot ? b->d_iq2 : b->d_iq;
    const long long row = (long long)n_blocks * FMRX_BLOCK_BYTES;
    // ingest: this slot's previous contents must have been consumed by the front end two submits ago
    if (b->iqfree_valid[slot]) CU(cudaStreamWaitEvent(b->s_in, b->e_iqfree[slot], 0));
    CU(cudaMemcpyAsync(dst, iq, (size_t)b->S * row, cudaMemcpyHostToDevice, b->s_in));
    CU(cudaEventRecord(b->e_h2d[slot], b->s_in));
    CU(cudaStreamWaitEvent(b->s_a, b->e_h2d[slot], 0));
    // the three-phase pipeline of the device-resident path
    const int set = (int)(b->calls % kSets);
    if (b->ev_c_valid[set]) CU(cudaStreamWaitEvent(b->s_a, b->ev_c[set], 0));
    if (b->d2h_valid) CU(cudaStreamWaitEvent(b->s_c, b->e_d2h, 0));  // phase C overwrites the result buffers the previous copy-out reads
    b->want_audio_f = o.audio_f != nullptr;
    if (int e = enqueue_chain(b, dst, row, 0, b->S, n_blocks, none, set, b->s_a, b->s_p, b->s_c)) return e;
    CU(cudaEventRecord(b->e_iqfree[slot], b->s_a));
    b->iqfree_valid[slot] = true;
    CU(cudaEventRecord(b->ev_c[set], b->s_c));
    b->ev_c_valid[set] = true;
    // egress
    CU(cudaStreamWaitEvent(b->s_out, b->ev_c[set], 0));
    if (int e = copy_outputs(b, 0, b->S, n_blocks, o, cudaMemcpyDeviceToHost, b->s_out)) return e;
    CU(cudaEventRecord(b->e_d2h, b->s_out));
    b->d2h_valid = true;
    *ticket = nsub;
    CU(cudaEventRecord(b->ev_ticket[nsub % kTickets], b->s_out));
    b->submits.store(nsub + 1);
    b->calls += 1;
    b->block_id += n_blocks;
    b->last_blocks = n_blocks;
    return FMRX_OK;
}

int fmrx_batch_wait(fmrx_batch *b, long long ticket) {
    if (!b) return fail(FMRX_ERR_ARG, "null handle");
    const long long issued = b->submits.load();
    if (ticket < 0 || ticket >= issued) return fail(FMRX_ERR_ARG, "ticket %lld was never issued (next is %lld)", ticket, issued);
    CU(cudaSetDevice(b->cfg.device));
    // the slot holds this ticket's event or, once recycled, that of a later submit on the same in-order stream
    CU(cudaEventSynchronize(b->ev_ticket[ticket % kTickets]));
    return FMRX_OK;
}

int fmrx_batch_wait_ingest(fmrx_batch *b, long long ticket) {
    if (!b) return fail(FMRX_ERR_ARG, "null handle");
    const long long issued = b->submits.load();
    if (ticket < 0 || ticket >= issued) return fail(FMRX_ERR_ARG, "ticket %lld was never issued (next is %lld)", ticket, issued);
    CU(cudaSetDevice(b->cfg.device));
    // the slot holds this ticket's copy event or, once reused, that of a later submit into the same staging buffer on the same
    // in-order stream (then the wait is longer than necessary, never shorter)
    CU(cudaEventSynchronize(b->e_h2d[ticket & 1]));
    return FMRX_OK;
}

int fmrx_batch_rds_offsets(fmrx_batch *b, int32_t *offsets) {
    if (!b || !offsets) return fail(FMRX_ERR_ARG, "null pointer");
    CU(cudaSetDevice(b->cfg.device));
    if (int e = fmrx_batch_sync(b)) return e;
    CU(cudaMemcpy2D(offsets, sizeof(int32_t), b->dec_st + 1, FMRX_RDS_STATE_WORDS * sizeof(int32_t), sizeof(int32_t), b->S, cudaMemcpyDeviceToHost));
    return FMRX_OK;
}

float fmrx_batch_rds_phase(const fmrx_batch *b) { return b ? b->rds_phase : 0.0f; }
int fmrx_batch_set_rds_phase(fmrx_batch *b, float phase_adj) {
    if (!b || !(phase_adj == phase_adj)) return fail(FMRX_ERR_ARG, "fmrx_batch_set_rds_phase: bad argument");
    b->rds_phase = phase_adj;
    return FMRX_OK;
}

int fmrx_batch_tap_len(const fmrx_batch *b, int which) {
    if (!b) return 0;
    switch (which) {
        case FMRX_TAP_MONO: case FMRX_TAP_STEREO: return b->n_audio;
        case FMRX_TAP_RDS_RES: case FMRX_TAP_RDS_RRC: return NRDS;
        default: return which >= 0 && which < FMRX_TAP_COUNT ? NIF : 0;
    }
}

int fmrx_batch_tap(fmrx_batch *b, int which, float *dst) {
    if (!b || !dst) return fail(FMRX_ERR_ARG, "null pointer");
    const float *src[FMRX_TAP_COUNT] = {b->demod, b->mono, b->pilot, b->nco, b->sbpf, b->stereo, b->rbpf, b->rsq, b->rnco, b->rlpf, b->rres, b->rrrc};
    if (which < 0 || which >= FMRX_TAP_COUNT || !src[which] || b->last_blocks == 0) return fail(FMRX_ERR_STATE, "tap %d not available", which);
    CU(cudaSetDevice(b->cfg.device));
    if (int e = fmrx_batch_sync(b)) return e;
    const size_t len = fmrx_batch_tap_len(b, which), w = len * b->last_blocks * 4;
    const bool dbl = which == FMRX_TAP_MONO || which == FMRX_TAP_PILOT || which == FMRX_TAP_NCO || which == FMRX_TAP_STEREO_BPF || which == FMRX_TAP_RDS_BPF ||
                     which == FMRX_TAP_RDS_SQ || which == FMRX_TAP_RDS_NCO;
    const float *base = src[which] + (dbl ? (size_t)b->last_set * (which == FMRX_TAP_MONO ? b->set_au : b->set_if) : 0);
    CU(cudaMemcpy2D(dst, w, base, len * b->NB * 4, w, b->S, cudaMemcpyDeviceToHost));
    return FMRX_OK;
}

// checkpoint blob = header + the device state blob.  The header pins everything the layout of the state depends on, so a
// blob from a handle of another shape, a truncated one or one written by another layout version is refused, not applied.
namespace {
struct StateHeader {
    uint32_t magic, version;      // 'FMRX' little endian, kStateLayout
    int32_t mode, profile, n_streams, paths, rds_fast, reserved;
    uint64_t state_bytes;
    long long block_id;
};
constexpr uint32_t kStateMagic = 0x58524D46u, kStateLayout = 4;  // bump kStateLayout when a segment or the decoder words (fmrx_rds.cu W_*) change
StateHeader make_header(const fmrx_batch *b) {
    StateHeader h{};
    h.magic = kStateMagic; h.version = kStateLayout; h.mode = b->cfg.mode; h.profile = b->cfg.profile; h.n_streams = b->S;
    h.paths = (b->audio_on ? FMRX_PATH_AUDIO : 0) | (b->rds_on ? FMRX_PATH_RDS : 0); h.rds_fast = b->rds_fast ? 1 : 0;
    h.state_bytes = b->state_bytes; h.block_id = b->block_id;
    return h;
}
}  // namespace

size_t fmrx_batch_state_bytes(const fmrx_batch *b) { return b ? b->state_bytes + sizeof(StateHeader) : 0; }

int fmrx_batch_get_state(fmrx_batch *b, void *blob, size_t bytes) {
    if (!b || !blob) return fail(FMRX_ERR_ARG, "null pointer");
    if (bytes < fmrx_batch_state_bytes(b)) return fail(FMRX_ERR_ARG, "state buffer of %zu bytes, %zu needed", bytes, fmrx_batch_state_bytes(b));
    CU(cudaSetDevice(b->cfg.device));
    if (int e = fmrx_batch_sync(b)) return e;
    const StateHeader h = make_header(b);
    std::memcpy(blob, &h, sizeof(h));
    CU(cudaMemcpy((char *)blob + sizeof(h), b->d_state, b->state_bytes, cudaMemcpyDeviceToHost));
    return FMRX_OK;
}

int fmrx_batch_set_state(fmrx_batch *b, const void *blob, size_t bytes) {
    if (!b || !blob) return fail(FMRX_ERR_ARG, "null pointer");
    StateHeader h{};
    if (bytes < sizeof(h)) return fail(FMRX_ERR_ARG, "state blob of %zu bytes is shorter than its header", bytes);
    std::memcpy(&h, blob, sizeof(h));
    const StateHeader want = make_header(b);
    if (h.magic != kStateMagic || h.version != kStateLayout) return fail(FMRX_ERR_ARG, "not a state blob of this library version (magic %08x, layout %u; expected layout %u)", h.magic, h.version, kStateLayout);
    if (h.mode != want.mode || h.profile != want.profile || h.n_streams != want.n_streams || h.paths != want.paths || h.rds_fast != want.rds_fast || h.state_bytes != want.state_bytes)  // the two RDS back ends keep different states
        return fail(FMRX_ERR_ARG, "state blob is from a handle of another shape (mode %d profile %d streams %d paths %d staged-RDS %d, %llu bytes; this handle: %d %d %d %d %d, %llu)",
                    h.mode, h.profile, h.n_streams, h.paths, !h.rds_fast, (unsigned long long)h.state_bytes, want.mode, want.profile, want.n_streams, want.paths, !want.rds_fast, (unsigned long long)want.state_bytes);
    if (bytes < sizeof(h) + b->state_bytes) return fail(FMRX_ERR_ARG, "state blob truncated: %zu bytes, %zu needed", bytes, sizeof(h) + b->state_bytes);
    if (h.block_id < 0) return fail(FMRX_ERR_ARG, "state blob carries a negative block counter");
    CU(cudaSetDevice(b->cfg.device));
    if (int e = fmrx_batch_sync(b)) return e;
    b->block_id = h.block_id;
    CU(cudaMemcpy(b->d_state, (const char *)blob + sizeof(h), b->state_bytes, cudaMemcpyHostToDevice));
    return FMRX_OK;
}

long long fmrx_batch_block_id(const fmrx_batch *b) { return b ? b->block_id : -1; }

int fmrx_batch_profile(fmrx_batch *b, int enable) {
    if (!b) return fail(FMRX_ERR_ARG, "null handle");
    if (int e = fmrx_batch_sync(b)) return e;
    b->profiling = enable != 0;
    b->profile_pipelined = enable == 2;
    b->marks.clear();
    b->prof_used = 0;
    for (int i = 0; i < FMRX_STAGE_COUNT; ++i) { b->stage_ms[i] = 0; b->stage_launches[i] = 0; }
    return FMRX_OK;
}

int fmrx_batch_stage_times(fmrx_batch *b, double *ms, long long *count) {
    if (!b || !ms) return fail(FMRX_ERR_ARG, "null pointer");
    if (int e = fmrx_batch_sync(b)) return e;
    for (auto &m : b->marks) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, m.t0, m.t1));
        b->stage_ms[m.stage] += t;
        b->stage_launches[m.stage] += 1;
    }
    b->marks.clear();
    b->prof_used = 0;
    for (int i = 0; i < FMRX_STAGE_COUNT; ++i) { ms[i] = b->stage_ms[i]; if (count) count[i] = b->stage_launches[i]; }
    return FMRX_OK;
}

int fmrx_batch_timeline(fmrx_batch *b, int cap, int32_t *stage, float *t0_ms, float *t1_ms) {
    if (!b || !stage || !t0_ms || !t1_ms) return fail(FMRX_ERR_ARG, "null pointer");
    if (int e = fmrx_batch_sync(b)) return e;
    int n = 0;
    for (auto &m : b->marks) {
        if (n >= cap) break;
        stage[n] = m.stage;
        CU(cudaEventElapsedTime(&t0_ms[n], b->marks[0].t0, m.t0));
        CU(cudaEventElapsedTime(&t1_ms[n], b->marks[0].t0, m.t1));
        ++n;
    }
    return n;
}

int fmrx_pinned_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(FMRX_ERR_ARG, "null pointer");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(FMRX_ERR_ALLOC, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    return FMRX_OK;
}

int fmrx_pinned_free(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return FMRX_OK;
}

int fmrx_measure_fp32_peak(int device, int kind, int reps, double *tera) {
    if (!tera) return fail(FMRX_ERR_ARG, "null pointer");
    CU(cudaSetDevice(device));
    LAUNCH(measure_fp32_peak(device, kind, reps, tera));
    return FMRX_OK;
}

}  // extern "C"
