// "Model-compatible" operators: the arithmetic of the reference's PYTHON models, in float64, so that
// model/fmMonoBlock.py / model/fmRDSblock.py can be diffed tightly against the GPU (the C++ program is numerically a
// different receiver: fp32, derivative discriminator, one-late FIR state, its own filter design -- SURVEY App. C).
//
//   fmrx_model_firwin    scipy.signal.firwin(N, cutoff, window='hann'[, pass_zero='bandpass']) as the models call it
//                        (model/fmMonoBlock.py:43-45,115,150,159; model/fmRDSblock.py:64-105): host code, float64.
//   fmrx_model_lfilter   scipy.signal.lfilter(b, 1.0, x, zi=state) followed by the models' [::decim] slicing, optionally on
//                        the zero-stuffed input of the RDS resampler (model/fmRDSblock.py:188-199).  lfilter is a
//                        transposed direct form II: for an FIR that is y[n] = b0 x[n] + (b1 x[n-1] + (... + b_{N-1}
//                        x[n-N+1])), products and sums rounded one by one, OLDEST tap first -- evaluated in that
//                        order; only the retained outputs are computed.  (scipy itself routes an FIR through
//                        numpy.convolve, whose summation order is an implementation detail, so agreement is to a few
//                        ulp of the accumulated magnitude: 1e-13 relative, tests/test_model_ops.py.)  The carried state
//                        is kept as the last N-1 input samples (the partial sums lfilter carries are a function of
//                        exactly those).
//   fmrx_model_demod     fmSupportLib.fmDemodArctan (model/fmSupportLib.py:12-44): atan2, numpy.unwrap against the
//                        carried UNWRAPPED phase, difference.  Sample-serial (the unwrapped phase is the state): one
//                        stream per lane.
//   fmrx_model_pll       fmPll.fmPll (model/fmPll.py:4-56), float64 throughout, in-phase and quadrature NCO outputs,
//                        the model's state order.  One stream per lane.
// These are parity tools, not the throughput path: one thread per output / per stream, no tiling.
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

// x_ext(p) for p in [-(nt-1), n): history then block, both in INPUT samples
__global__ void model_lfilter_kernel(double *y, const double *x, const double *b, const double *hist, int n, int nt, int ny, int decim, int up, int n_streams) {
    const int s = blockIdx.y;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= ny) return;
    const double *xs = x + (long long)s * n;
    const double *hs = hist + (long long)s * (nt - 1);
    const long long pos = (long long)decim * m;  // index of this output in the (zero-stuffed) filter-rate stream
    double acc = 0.0;
    for (int k = nt - 1; k >= 0; --k) {          // oldest tap first: the nesting order of a transposed direct form II
        const long long p = pos - k;
        double v = 0.0;
        if (up == 1 || (p % up) == 0) {           // a stuffed zero adds b*0 = +-0 to the sum: no effect, skipped
            // p may be negative: floor division by `up` of a multiple of `up` is exact either way
            const long long q = up == 1 ? p : p / up;
            if (q >= 0) v = xs[q];
            else if (q >= -(long long)(nt - 1)) v = hs[nt - 1 + q];
            acc = __dadd_rn(__dmul_rn(b[k], v), acc);
        }
    }
    y[(long long)s * ny + m] = acc;
}

__global__ void model_hist_kernel(double *hist, const double *x, int n, int nt, int n_streams) {
    // new history = last L samples of [old history, x]; one block per stream, every read before any write (n may be < L)
    const int s = blockIdx.x, L = nt - 1;
    double v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + j * 1024;
        if (i >= L) continue;
        const int q = n - L + i;
        v[j] = q >= 0 ? x[(long long)s * n + q] : hist[(long long)s * L + (L + q)];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + j * 1024;
        if (i < L) hist[(long long)s * L + i] = v[j];
    }
}

// numpy.mod for float64 (floor modulo, result takes the divisor's sign), divisor > 0
__device__ __forceinline__ double np_mod(double a, double b) {
    double r = fmod(a, b);
    if (r != 0.0 && r < 0.0) r += b;
    return r;
}

__global__ void model_demod_kernel(double *out, const double *I, const double *Q, int n, int n_streams, double *prev_phase) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    double prev = prev_phase[s];
    const double *is = I + (long long)s * n, *qs = Q + (long long)s * n;
    double *os = out + (long long)s * n;
    for (int k = 0; k < n; ++k) {
        double cur = atan2(qs[k], is[k]);
        // numpy.unwrap([prev, cur]) (numpy/lib/function_base.py): period 2*pi, discont pi
        const double dd = cur - prev;
        double ddmod = np_mod(dd + PI, TWO_PI) - PI;
        if (ddmod == -PI && dd > 0.0) ddmod = PI;
        double corr = ddmod - dd;
        if (fabs(dd) < PI) corr = 0.0;
        cur = cur + corr;
        os[k] = cur - prev;
        prev = cur;
    }
    prev_phase[s] = prev;
}

__global__ void model_pll_kernel(double *nco, double *ncoq, const double *x, int n, int n_streams, double freq, double Fs, double scale, double adj, double bw, double *state) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const double Cp = 2.666, Ci = 3.555;
    const double Kp = bw * Cp, Ki = (bw * bw) * Ci;  // model/fmPll.py:7-10 (same association)
    double *st = state + (long long)s * 6;
    double integ = st[0], phase = st[1], fbi = st[2], fbq = st[3];
    const double nco0 = st[4], off = st[5];
    const double *xs = x + (long long)s * n;
    double *oi = nco + (long long)s * (n + 1), *oq = ncoq + (long long)s * (n + 1);
    oi[0] = nco0;
    oq[0] = 0.0;  // the model leaves ncoOutQ[0] uninitialised (np.empty); defined as 0 here
    const double w = __dmul_rn(__dmul_rn(2.0, 3.141592653589793), freq / Fs);  // 2*math.pi*(freq/Fs)
    for (int k = 0; k < n; ++k) {
        const double eI = __dmul_rn(xs[k], fbi), eQ = __dmul_rn(xs[k], -fbq);
        const double eD = atan2(eQ, eI);
        integ = __dadd_rn(integ, __dmul_rn(Ki, eD));
        phase = __dadd_rn(__dadd_rn(phase, __dmul_rn(Kp, eD)), integ);
        const double trig = __dadd_rn(__dmul_rn(w, __dadd_rn(__dadd_rn(off, (double)k), 1.0)), phase);
        fbi = cos(trig);
        fbq = sin(trig);
        const double targ = __dadd_rn(__dmul_rn(trig, scale), adj);
        oi[k + 1] = cos(targ);
        oq[k + 1] = sin(targ);
    }
    st[0] = integ; st[1] = phase; st[2] = fbi; st[3] = fbq; st[4] = oi[n]; st[5] = off + (double)n;
}

template <class T>
struct DBuf {
    T *p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
    cudaError_t up(const T *h, size_t n) { cudaError_t e = alloc(n); return e ? e : cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T *h, size_t n) { return cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};

#define MCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fmrx::fail(FMRX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

}  // namespace
}  // namespace fmrx

extern "C" {

int fmrx_model_firwin(int ntaps, const double *cutoff, int n_cutoff, int pass_zero, double *h) {
    if (ntaps <= 0 || !cutoff || !h || n_cutoff <= 0 || n_cutoff > 2) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_firwin: bad argument");
    for (int i = 0; i < n_cutoff; ++i)
        if (!(cutoff[i] > 0.0 && cutoff[i] < 1.0) || (i && cutoff[i] <= cutoff[i - 1])) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_firwin: cutoffs must increase inside (0, 1)");
    if ((n_cutoff == 1) != (pass_zero != 0)) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_firwin: one cutoff = low-pass (pass_zero 1), two = band-pass (pass_zero 0)");
    if (!pass_zero && ntaps % 2 == 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_firwin: a band-pass design needs an odd number of taps (scipy raises too)");
    // scipy/signal/_fir_filter_design.py firwin: bands from the cutoffs, h = sum(right*sinc(right*m) - left*sinc(left*m)),
    // symmetric Hann window, scaled to unity gain at the centre of the first pass band
    const double PI = 3.141592653589793;
    const double left = pass_zero ? 0.0 : cutoff[0], right = pass_zero ? cutoff[0] : cutoff[1];
    const double alpha = 0.5 * (ntaps - 1);
    auto sinc = [&](double v) { return v == 0.0 ? 1.0 : std::sin(PI * v) / (PI * v); };  // numpy.sinc
    const double scale_frequency = left == 0.0 ? 0.0 : 0.5 * (left + right);
    double sum = 0.0;
    for (int i = 0; i < ntaps; ++i) {
        const double m = i - alpha;
        double v = right * sinc(right * m) - left * sinc(left * m);
        const double win = ntaps == 1 ? 1.0 : 0.5 - 0.5 * std::cos(2.0 * PI * i / (ntaps - 1));
        v *= win;
        h[i] = v;
        sum += v * std::cos(PI * m * scale_frequency);
    }
    for (int i = 0; i < ntaps; ++i) h[i] /= sum;
    return FMRX_OK;
}

int fmrx_model_lfilter(double *y, const double *x, int n_streams, int n, const double *b, int ntaps, double *hist, int decim, int up) {
    if (!y || !x || !b || !hist || n_streams <= 0 || n <= 0 || ntaps < 2 || ntaps > 4097 || decim <= 0 || up <= 0)
        return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_lfilter: bad argument");
    if (((long long)n * up) % decim != 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_lfilter: n*up (%lld) must be a multiple of decim (%d): the models slice whole blocks", (long long)n * up, decim);
    const int ny = (int)(((long long)n * up) / decim);
    const size_t nx = (size_t)n_streams * n, nh = (size_t)n_streams * (ntaps - 1);
    fmrx::DBuf<double> dx, dy, db, dh;
    MCU(dx.up(x, nx)); MCU(dy.alloc((size_t)n_streams * ny)); MCU(db.up(b, ntaps)); MCU(dh.up(hist, nh));
    dim3 grid((ny + 127) / 128, n_streams);
    fmrx::model_lfilter_kernel<<<grid, 128>>>(dy.p, dx.p, db.p, dh.p, n, ntaps, ny, decim, up, n_streams);
    MCU(cudaGetLastError());
    fmrx::model_hist_kernel<<<n_streams, 1024>>>(dh.p, dx.p, n, ntaps, n_streams);
    fmrx::launch_counter() += 2;
    MCU(cudaGetLastError());
    MCU(dy.down(y, (size_t)n_streams * ny)); MCU(dh.down(hist, nh));
    return FMRX_OK;
}

int fmrx_model_demod(double *demod, const double *I, const double *Q, int n_streams, int n, double *prev_phase) {
    if (!demod || !I || !Q || !prev_phase || n_streams <= 0 || n <= 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_demod: bad argument");
    const size_t nx = (size_t)n_streams * n;
    fmrx::DBuf<double> di, dq, dout, dp;
    MCU(di.up(I, nx)); MCU(dq.up(Q, nx)); MCU(dout.alloc(nx)); MCU(dp.up(prev_phase, n_streams));
    fmrx::model_demod_kernel<<<(n_streams + 31) / 32, 32>>>(dout.p, di.p, dq.p, n, n_streams, dp.p);
    fmrx::launch_counter() += 1;
    MCU(cudaGetLastError());
    MCU(dout.down(demod, nx)); MCU(dp.down(prev_phase, n_streams));
    return FMRX_OK;
}

int fmrx_model_pll(double *nco, double *nco_q, const double *x, int n_streams, int n, double freq, double Fs, double nco_scale, double phase_adjust,
                   double norm_bandwidth, double *state) {
    if (!nco || !nco_q || !x || !state || n_streams <= 0 || n <= 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_model_pll: bad argument");
    const size_t nx = (size_t)n_streams * n, no = (size_t)n_streams * (n + 1);
    fmrx::DBuf<double> dx, di, dq, ds;
    MCU(dx.up(x, nx)); MCU(di.alloc(no)); MCU(dq.alloc(no)); MCU(ds.up(state, (size_t)n_streams * 6));
    fmrx::model_pll_kernel<<<(n_streams + 31) / 32, 32>>>(di.p, dq.p, dx.p, n, n_streams, freq, Fs, nco_scale, phase_adjust, norm_bandwidth, ds.p);
    fmrx::launch_counter() += 1;
    MCU(cudaGetLastError());
    MCU(di.down(nco, no)); MCU(dq.down(nco_q, no)); MCU(ds.down(state, (size_t)n_streams * 6));
    return FMRX_OK;
}

}  // extern "C"
