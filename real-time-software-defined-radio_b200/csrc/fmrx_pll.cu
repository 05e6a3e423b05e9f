// Carrier-recovery PLL + NCO (SURVEY 8a rows a13, a14's loop part) and the L/R combiner + quantiser (row a16).
//
// Reference: fmPLL, /root/reference/src/helper.cpp:13-57 (identical loop body inside pllCombine, :149-159); state struct
// src/helper.h:17-19; call sites src/fm_radio.cpp:262 (19 kHz pilot, NCO x2) and :400 (114 kHz, NCO x0.5).
//
// The recurrence is sample-serial, so one stream runs per LANE and independent streams are packed 32 to a warp, one
// warp per CTA so the (few) warps spread over all SMs.  Types follow the reference exactly (App. D.3): loop state is
// fp32 with one rounding per operation (no FMA contraction), the phase detector and the oscillator are evaluated in
// double precision on float-valued arguments, and the oscillator argument is a double expression rounded to fp32.
// Bit-exactness matters here: once trigArg's fp32 ulp is coarse (after ~1 s of signal) two trajectories that differ in
// the last bit decorrelate at the ulp level, far above the 1e-5 audio tolerance.
//
// The kernel is bound by the LATENCY of that per-sample chain (8192 lanes = 256 warps on 592 schedulers), so the
// step is built from the short-chain routines of fmrx_pllmath.h instead of libm's atan2/sincos/cos: ~30 dependent
// double-precision operations per sample instead of ~150, and none of them on libm's large-argument slow path
// (the 114 kHz loop's argument passes 1e5 rad within three blocks).
#include <cuda_runtime.h>

#include <cstdlib>

#include "fmrx_internal.h"
#include "fmrx_pllmath.h"

namespace fmrx {
namespace {

using pllmath::PllCarry;
using pllmath::PllCoef;
using pllmath::PllFast;
using pllmath::PllLibmOut;
using pllmath::pll_step_fast;
using pllmath::pll_step_libm;

constexpr double kTwoPi = 2 * 3.14159265358979323846;  // `2*PI`, src/helper.cpp:41 with src/dy4.h:13

// (trigOffset + k) + 1 in fp32, as src/helper.cpp:41 forms it
__device__ __forceinline__ float count(float off, int k) { return __fadd_rn(__fadd_rn(off, (float)k), 1.0f); }

struct PllSide {
    const float *x;
    float *nco;
    float *state;
    const float *mul;  // optional second signal ...
    float *prod;       // ... and where nco[k] * mul[k] goes
    float Ki, Kp, fratio, scale, adj;
};

// V selects the step of fmrx_pllmath.h: 0 = conversions as instructions, both signs' theta0 prepared; 1 = integer-built
// conversions and theta0 for the next sample's sign only (fewer instructions on the FP64 and conversion pipes); 2 = as 1 with
// the two widenings that are off the dependency chain left as conversion instructions.
template <int V>
__global__ void __launch_bounds__(32, 8) pll_kernel(PllSide A, PllSide B, long long ld, int n_streams, int n, int n_blocks) {
    __shared__ __align__(16) pllmath::PllTheta theta[16];  // V = 1: (k*H1, k*L1) by (quadrant, sign of r, sign of the next sample)
    __shared__ double kdoubles[pllmath::kPllKDoubles];      // V = 1: the step's double constants, loaded so that they stay in registers
    if (V >= 1) {
        if (threadIdx.x < 16) theta[threadIdx.x] = pllmath::pll_theta_entry(threadIdx.x);
        if (threadIdx.x < pllmath::kPllKDoubles) kdoubles[threadIdx.x] = pllmath::pll_k_value(threadIdx.x);
        __syncthreads();
    }
    const pllmath::PllK K = V >= 1 ? pllmath::pll_k_from(kdoubles, 0x38000000u) : pllmath::PllK{};
    int lane = blockIdx.x * blockDim.x + threadIdx.x;
    const PllSide &P = lane < n_streams ? A : B;
    if (lane >= n_streams) lane -= n_streams;
    if (lane >= n_streams || P.x == nullptr) return;
    const float *x = P.x + (long long)lane * ld;
    float *nco = P.nco + (long long)lane * ld;
    const float *mul = P.mul ? P.mul + (long long)lane * ld : nullptr;
    float *prod = P.prod ? P.prod + (long long)lane * ld : nullptr;
    float *st = P.state + (long long)lane * 6;
    PllCarry c{st[0], st[1], st[2], st[3]};
    PllFast f;
    pllmath::pll_disarm(f);  // only the float state is carried between launches: the first group goes through libm
    float off = st[4];
    float last = st[5];
    const PllCoef p{P.Ki, P.Kp, P.scale, P.adj, __dmul_rn(kTwoPi, (double)P.fratio)};
    // the libm path, out of line: one step, everything by value
    auto slow = [&](float xin, float cnt, float xnext) {
        const PllLibmOut o = pll_step_libm(c, p, xin, cnt);
        c = o.c;
        if (V >= 1) pllmath::pll_rearm1(f, o.trig, xnext < 0.0f, theta, K);
        else pllmath::pll_rearm(f, o.trig);
        return o.nco;
    };
    for (int b = 0; b < n_blocks; ++b) {
        const float *xb = x + (long long)b * n;
        float *ob = nco + (long long)b * n;
        const float *mb = mul ? mul + (long long)b * n : nullptr;
        float *pb = prod ? prod + (long long)b * n : nullptr;
        int k = 0;
        const bool off_ok = V == 0 || (off >= 0.0f && off < 1.0e30f);  // V = 1 rebuilds (double)cnt from the bits of a normal float >= 1
        if ((((uintptr_t)xb | (uintptr_t)ob | (uintptr_t)mb | (uintptr_t)pb) & 15) == 0 && n >= 8) {
            // the input is read two groups (8 samples, ~2.5k cycles of loop time) ahead of its use: a lane streams its
            // own row, so every load is a cache line of its own and would otherwise sit on the dependency chain
            const float4 *x4 = reinterpret_cast<const float4 *>(xb);
            const int groups = n / 4;
            const float4 *m4 = reinterpret_cast<const float4 *>(mb);
            float4 v0 = __ldg(x4), v1 = __ldg(x4 + 1);
            const float x_after = (V >= 1 && b + 1 < n_blocks) ? __ldg(xb + n) : 1.0f;  // what follows this block's last sample
            float4 w0 = make_float4(0.f, 0.f, 0.f, 0.f), w1 = w0;
            if (pb) { w0 = __ldg(m4); w1 = __ldg(m4 + 1); }
            int g = 0;
            while (g < groups) {
                float4 v, w, o;
                float xn;
                bool redo = false;
                // The fast run.  No call inside this loop, so the double constants of the step stay in registers from one
                // iteration to the next (with the libm redo inside the loop ptxas rebuilt 35 of them at the top of every
                // iteration: 9 of 176 instructions per step).
                for (; g < groups; ++g, k += 4) {
                    v = v0;
                    w = w0;
                    v0 = v1;
                    w0 = w1;
                    if (g + 2 < groups) {
                        v1 = __ldg(x4 + g + 2);
                        if (pb) w1 = __ldg(m4 + g + 2);
                    }
                    // ... and its 128-byte line is pulled into L2 sixteen groups before that: with 8192 lanes each on a row of
                    // its own, every line is a DRAM page (and mostly a TLB) miss, and the register prefetch alone left 15% of
                    // the kernel's samples waiting on it (ncu source page, profiles/r1w)
                    if ((g & 7) == 0 && g + 16 < groups) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(x4 + g + 16));
                        if (pb) asm volatile("prefetch.global.L2 [%0];" ::"l"(m4 + g + 16));
                    }
                    // four branch-free steps = one basic block; if any of them left the fast path's domain (first group
                    // after loading the state, zero / non-finite input, the +-pi seam: ~1e-5 of the groups) the four
                    // carried floats are restored and the group is redone with libm
                    const PllCarry saved = c;
                    const float s_last = last;
                    bool ok0, ok1, ok2, ok3;
                    o.x = last;
                    xn = g + 1 < groups ? v0.x : x_after;  // the sample after this group (V = 1 prepares for its sign)
                    if (V >= 1) {
                        o.y = pllmath::pll_step_fast1<V == 2>(c, f, p, K, v.x, count(off, k), v.y < 0.0f, theta, ok0);
                        o.z = pllmath::pll_step_fast1<V == 2>(c, f, p, K, v.y, count(off, k + 1), v.z < 0.0f, theta, ok1);
                        o.w = pllmath::pll_step_fast1<V == 2>(c, f, p, K, v.z, count(off, k + 2), v.w < 0.0f, theta, ok2);
                        last = pllmath::pll_step_fast1<V == 2>(c, f, p, K, v.w, count(off, k + 3), xn < 0.0f, theta, ok3);
                    } else {
                        o.y = pll_step_fast(c, f, p, v.x, count(off, k), ok0);
                        o.z = pll_step_fast(c, f, p, v.y, count(off, k + 1), ok1);
                        o.w = pll_step_fast(c, f, p, v.z, count(off, k + 2), ok2);
                        last = pll_step_fast(c, f, p, v.w, count(off, k + 3), ok3);
                    }
                    if (!(ok0 && ok1 && ok2 && ok3 && off_ok)) {
                        c = saved;
                        last = s_last;
                        redo = true;
                        break;
                    }
                    *reinterpret_cast<float4 *>(ob + k) = o;
                    if (pb) *reinterpret_cast<float4 *>(pb + k) = make_float4(__fmul_rn(w.x, o.x), __fmul_rn(w.y, o.y), __fmul_rn(w.z, o.z), __fmul_rn(w.w, o.w));
                }
                if (!redo) break;
                o.x = last;
                o.y = slow(v.x, count(off, k), v.y);
                o.z = slow(v.y, count(off, k + 1), v.z);
                o.w = slow(v.z, count(off, k + 2), v.w);
                last = slow(v.w, count(off, k + 3), xn);
                *reinterpret_cast<float4 *>(ob + k) = o;
                if (pb) *reinterpret_cast<float4 *>(pb + k) = make_float4(__fmul_rn(w.x, o.x), __fmul_rn(w.y, o.y), __fmul_rn(w.z, o.z), __fmul_rn(w.w, o.w));
                ++g;
                k += 4;
            }
        }
        for (; k < n; ++k) {
            ob[k] = last;  // output sample k is the NCO value of step k-1 (src/helper.cpp:29,44,56)
            if (pb) pb[k] = __fmul_rn(mb[k], last);
            last = slow(xb[k], count(off, k), (k + 1 < n || b + 1 < n_blocks) ? xb[k + 1] : 1.0f);
        }
        off = __fadd_rn(off, (float)n);  // src/helper.cpp:53
    }
    st[0] = c.integ; st[1] = c.phase; st[2] = c.fbi; st[3] = c.fbq; st[4] = off; st[5] = last;
}

// src/fm_radio.cpp:277-299: L=(m+s)/2, R=(m-s)/2; NaN -> 0 else static_cast<short>(x*16384*mult), which on x86-64 is
// cvttss2si (truncate toward zero, 0x80000000 when out of range) followed by taking the low 16 bits.
__device__ __forceinline__ int16_t quantise(float v, float mult) {
    if (isnan(v)) return 0;
    const float s = __fmul_rn(__fmul_rn(v, 16384.0f), mult);
    const int w = (s > -2147483648.0f && s < 2147483648.0f) ? __float2int_rz(s) : (int)0x80000000;
    return (int16_t)(unsigned short)(w & 0xFFFF);
}

// mono_delay > 0 (quality profile): the mono sum is taken `mono_delay` samples late, from the previous samples of this call or,
// at the start of the call, from the carried tail of the previous one -- the L-R branch passes one more 151-tap filter (the 22-54 kHz
// band-pass, 75 samples at 240 kHz = 15 at 48 kHz) than the L+R branch, which the reference does not compensate
__global__ void combine_kernel(const float *mono, const float *stereo, int16_t *audio, float *audio_f, long long ld, int n_total, float mult, int mono_delay,
                               const float *mono_tail) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const float m = mono_delay == 0 ? mono[(long long)s * ld + i] : (i >= mono_delay ? mono[(long long)s * ld + i - mono_delay] : mono_tail[s * 16 + i]);
    const float t = stereo ? stereo[(long long)s * ld + i] : 0.0f;
    const float l = __fmul_rn(__fadd_rn(m, t), 0.5f), r = __fmul_rn(__fsub_rn(m, t), 0.5f);  // the `/2` of src/fm_radio.cpp:250-251: the same correctly rounded value, without the division sequence
    const long long o = ((long long)s * ld + i) * 2;
    if (audio_f) *reinterpret_cast<float2 *>(audio_f + o) = make_float2(l, r);
    if (audio) *reinterpret_cast<short2 *>(audio + o) = make_short2(quantise(l, mult), quantise(r, mult));
}

// the same for four consecutive samples per thread (128-bit loads of both inputs, one 128-bit store of the four L/R int16 pairs):
// the combiner moves 36 bytes per output quad and was bound by the number of requests in flight, not by arithmetic
__global__ void combine4_kernel(const float *mono, const float *stereo, int16_t *audio, float *audio_f, long long ld, int n_quads, float mult) {
    const int s = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads) return;
    const long long at = (long long)s * ld + 4LL * q;
    const float4 m = __ldg(reinterpret_cast<const float4 *>(mono + at));
    const float4 t = stereo ? __ldg(reinterpret_cast<const float4 *>(stereo + at)) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float mv[4] = {m.x, m.y, m.z, m.w}, tv[4] = {t.x, t.y, t.z, t.w};
    float l[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { l[i] = __fmul_rn(__fadd_rn(mv[i], tv[i]), 0.5f); r[i] = __fmul_rn(__fsub_rn(mv[i], tv[i]), 0.5f); }
    if (audio_f) {
        float4 *f = reinterpret_cast<float4 *>(audio_f + 2 * at);
        f[0] = make_float4(l[0], r[0], l[1], r[1]);
        f[1] = make_float4(l[2], r[2], l[3], r[3]);
    }
    if (audio) {
        auto pack = [&](int i) { return (unsigned)(unsigned short)quantise(l[i], mult) | ((unsigned)(unsigned short)quantise(r[i], mult) << 16); };
        *reinterpret_cast<uint4 *>(audio + 2 * at) = make_uint4(pack(0), pack(1), pack(2), pack(3));
    }
}

// after the combiner has read the old tail: the last `mono_delay` mono samples of this call become the next call's tail
__global__ void mono_tail_kernel(const float *mono, float *mono_tail, long long ld, int n_total, int mono_delay, int n_streams) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = g / 16, j = g % 16;
    if (s >= n_streams || j >= mono_delay) return;
    const int i = n_total - mono_delay + j;
    mono_tail[s * 16 + j] = i >= 0 ? mono[(long long)s * ld + i] : mono_tail[s * 16 + j + n_total];  // a call shorter than the delay shifts the tail (never in the chain)
}

// ------------------------------------------------------------------------------------------------------------------
// De-emphasis (quality profile; the reference sends the low-pass output straight to the quantiser, src/fm_radio.cpp:277-299):
// 1 / (1 + s tau) through the bilinear transform, y[n] = b (x[n] + x[n-1]) - a1 y[n-1], on L and R, then the quantiser.
// A first-order recursion is an affine map per sample, y -> c y + u, and affine maps compose: one CTA per station, one warp
// per channel, a lane owns a run of consecutive samples, runs it once from y = 0 to get its map (c^len, end value), the 32 maps
// are combined by a warp scan, and the run is redone from the right starting value.  The block is staged in shared memory so that
// global loads and stores stay coalesced.  State per station: x[-1], y[-1] of L and of R.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) deemphasis_kernel(float *audio_f, int16_t *audio, long long ld2, int n, int n_blocks, float b, float c, float mult,
                                                         float *state, int write_float) {
    extern __shared__ float blk[];  // 2n floats: L,R interleaved
    const int s = blockIdx.x, lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    float xprev = state[s * 4 + 2 * ch], yprev = state[s * 4 + 2 * ch + 1];
    const int len = (n + 31) / 32, lo = min(n, lane * len), hi = min(n, lo + len);
    for (int bk = 0; bk < n_blocks; ++bk) {
        float *g = audio_f + (long long)s * ld2 + (long long)bk * 2 * n;
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += 64) { const float v = g[i]; blk[i] = isnan(v) ? 0.0f : v; }  // the quantiser writes 0 for a NaN (mode 1, Q5)
        __syncthreads();
        const float x0 = lo > 0 ? blk[2 * (lo - 1) + ch] : xprev;
        float y = 0.0f, a = 1.0f, xp = x0;
        for (int i = lo; i < hi; ++i) {
            const float x = blk[2 * i + ch];
            y = fmaf(c, y, __fmul_rn(b, __fadd_rn(x, xp)));
            a *= c;
            xp = x;
        }
        // inclusive scan of the maps (A, B): applying lane j's map after the maps of the lanes below it
        float A = a, B = y;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float Ad = __shfl_up_sync(0xffffffffu, A, d), Bd = __shfl_up_sync(0xffffffffu, B, d);
            if (lane >= d) { B = fmaf(A, Bd, B); A *= Ad; }
        }
        float Ain = __shfl_up_sync(0xffffffffu, A, 1), Bin = __shfl_up_sync(0xffffffffu, B, 1);
        if (lane == 0) { Ain = 1.0f; Bin = 0.0f; }
        y = fmaf(Ain, yprev, Bin);  // the value just before this lane's run
        xp = x0;
        const float x_last = blk[2 * (n - 1) + ch];
        __syncwarp();
        for (int i = lo; i < hi; ++i) {
            const float x = blk[2 * i + ch];
            y = fmaf(c, y, __fmul_rn(b, __fadd_rn(x, xp)));
            xp = x;
            blk[2 * i + ch] = y;
        }
        yprev = __shfl_sync(0xffffffffu, fmaf(A, yprev, B), 31);  // the whole block's map applied to the carried value = y[n-1]
        xprev = x_last;
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += 64) {
            const float v = blk[i];
            if (write_float) g[i] = v;
            if (audio) audio[(long long)s * ld2 + (long long)bk * 2 * n + i] = quantise(v, mult);
        }
    }
    if (lane == 0) { state[s * 4 + 2 * ch] = xprev; state[s * 4 + 2 * ch + 1] = yprev; }
}

__global__ void multiply_kernel(const float *a, const float *b, float *y, long long ld, int n_total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const long long q = (long long)blockIdx.y * ld + i;
    y[q] = __fmul_rn(a[q], b[q]);
}

PllSide make_side(const float *x, float *nco, float *state, const PllParams &p, const float *mul, float *prod) {
    PllSide s;
    s.x = x; s.nco = nco; s.state = state;
    s.mul = (mul && prod) ? mul : nullptr; s.prod = (mul && prod) ? prod : nullptr;
    const float Cp = 2.666f, Ci = 3.555f;  // src/helper.cpp:15-16
    s.Ki = (p.bw * p.bw) * Ci;
    s.Kp = p.bw * Cp;
    s.fratio = p.freq / p.Fs;
    s.scale = p.scale;
    s.adj = p.phase_adj;
    return s;
}

}  // namespace

int launch_pll_blocks(const float *xa, float *ncoa, PllParams pa, float *sta, const float *xb, float *ncob, PllParams pb, float *stb,
                      long long ld, int n_streams, int n, int n_blocks, fmrx_stream_t st, const float *mula, float *proda, const float *mulb, float *prodb) {
    PllSide A = make_side(xa, ncoa, sta, pa, mula, proda);
    PllSide B = xb ? make_side(xb, ncob, stb, pb, mulb, prodb) : PllSide{};
    const int lanes = xb ? 2 * n_streams : n_streams;
    static const int variant = [] { const char *e = getenv("FMRX_PLL_STEP"); return e ? atoi(e) : 2; }();
    const int cta = 32, grid = (lanes + cta - 1) / cta;  // one warp per CTA (64- and 128-thread CTAs measure the same: the block scheduler already spreads the warps over the schedulers)
    if (variant == 0) pll_kernel<0><<<grid, cta, 0, st>>>(A, B, ld, n_streams, n, n_blocks);
    else if (variant == 2) pll_kernel<2><<<grid, cta, 0, st>>>(A, B, ld, n_streams, n, n_blocks);
    else pll_kernel<1><<<grid, cta, 0, st>>>(A, B, ld, n_streams, n, n_blocks);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

int launch_multiply(const float *a, const float *b, float *y, long long ld, int n_total, int n_streams, fmrx_stream_t st) {
    dim3 grid((n_total + 255) / 256, n_streams);
    multiply_kernel<<<grid, 256, 0, st>>>(a, b, y, ld, n_total);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

int launch_combine(const CombineJob &j, fmrx_stream_t st) {
    dim3 grid((j.n_total + 255) / 256, j.n_streams);
    const int dly = j.mono_tail ? j.mono_delay : 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(j.mono) | reinterpret_cast<uintptr_t>(j.stereo) | reinterpret_cast<uintptr_t>(j.audio) | reinterpret_cast<uintptr_t>(j.audio_f)) & 15) == 0;
    if (dly == 0 && aligned && j.n_total % 4 == 0 && j.ld % 4 == 0) {
        const int nq = j.n_total / 4;
        combine4_kernel<<<dim3((nq + 255) / 256, j.n_streams), 256, 0, st>>>(j.mono, j.stereo, j.audio, j.audio_f, j.ld, nq, (float)j.mult);
        launch_counter() += 1;
        return (int)cudaGetLastError();
    }
    combine_kernel<<<grid, 256, 0, st>>>(j.mono, j.stereo, j.audio, j.audio_f, j.ld, j.n_total, (float)j.mult, dly, j.mono_tail);
    launch_counter() += 1;
    if (dly > 0) {
        mono_tail_kernel<<<(j.n_streams * 16 + 255) / 256, 256, 0, st>>>(j.mono, j.mono_tail, j.ld, j.n_total, dly, j.n_streams);
        launch_counter() += 1;
    }
    return (int)cudaGetLastError();
}

int launch_deemphasis(float *audio_f, int16_t *audio, long long ld2, int n, int n_blocks, int n_streams, double b, double a1, int mult, float *state, int write_float,
                      fmrx_stream_t st) {
    const size_t smem = (size_t)2 * n * sizeof(float);
    if (smem > 48 * 1024) {
        if (cudaError_t e = cudaFuncSetAttribute(deemphasis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) return (int)e;
    }
    deemphasis_kernel<<<n_streams, 64, smem, st>>>(audio_f, audio, ld2, n, n_blocks, (float)b, (float)(-a1), (float)mult, state, write_float);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
