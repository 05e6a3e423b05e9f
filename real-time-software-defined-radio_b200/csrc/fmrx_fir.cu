// Stateful 151-tap FIR / decimator kernels (SURVEY 8a rows a6, a8, a9, a14's filter part, a15, a16's mixer) and the
// fused RF front end (rows a1, a2, a6, a7).
//
// Reference semantics reproduced here (file:line under /root/reference):
//   src/filter.cpp:126-219   y[n] = sum_k h[k] * X(d*n - k), k ascending, fp32 multiply then fp32 add (no FMA in the
//                            reference build); X(p>=0) = x[p]; X(-j) = zi[Z-j]; afterwards zi[i] = x[N-Z-1+i], i.e. the
//                            history is one sample late (Q1).  Inside a multi-block launch the history of block b>0
//                            is read straight from block b-1 of the input at the shifted index, so no state round trip.
//   src/helper.cpp:139,162   the squared-input variant;  src/filter.cpp:387,399 the mixer variant with half-weight history.
//   src/iofunc.cpp:67, src/fm_radio.cpp:68-72, src/rf_module.cpp:13-34 for the front end.
//
// B200 mapping: one CTA = NT threads x R=8 consecutive outputs = one tile of one (stream, block).  The input span is
// staged once in shared memory, 128 bits at a time, with a row pitch of d*R+4 words: lane t's window starts at
// (d*R+4)*t, which keeps every quad 16-byte aligned and makes the 128-bit loads of a quarter-warp cover all 32 banks
// ((d*R+4)/4 is odd), so both the staging stores and the window loads are conflict-free (the first version used
// scalar stores with a pitch of d*R+1: ncu showed a 4-way conflict on every staging store and 38% of the warp stalls
// on the shared-memory scoreboard, profiles/r1k_*).  The inner loop runs over the thread's INPUT samples, newest
// first: each sample is read from shared memory exactly once (one LDS.128 per four) and applied to every output it
// contributes to, which for a fixed output still visits the taps in ascending order -- the reference's summation order.  The loop is fully unrolled; taps arrive BY VALUE in the kernel parameter
// block, so each tap is a constant-bank operand of the multiply and the instruction stream is ~95% FMUL/FADD (or FFMA).
// `EXACT` keeps the reference's two roundings per tap (FMUL, FADD); otherwise FFMA.  The packed f32x2 forms are not
// used: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (checked with cuobjdump), which breaks the two
// roundings, and FFMA2 has the same lane throughput as FFMA anyway (measured, bench.py peak kinds 0 and 2).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

constexpr int R = 8;          // outputs per thread
constexpr int CT = 128;       // threads per CTA, single-channel kernels
constexpr int CTQ = 64;       // threads per CTA, two-channel kernels (float2 staging: 42 KB per 512-output tile, 5 CTAs/SM)
constexpr int OFF = 168;      // tile origin sits OFF input samples before the first output's newest sample: >= 150 + 10
                              // (history of the output just before the tile, for the discriminator) and 2*OFF % 16 == 0

struct Taps {
    float h[kTaps + 1];
};

template <int D, int NT>
struct Geom {
    static constexpr int TO = R * NT;          // outputs per tile
    static constexpr int ROW = D * R;          // input samples between consecutive threads' windows
    static constexpr int PITCH = ROW + 1;      // padded row -> odd lane stride
    static constexpr int SPAN = D * (TO - 1) + OFF + 1;
    static constexpr int WORDS = SPAN + SPAN / ROW + 1;
    static constexpr int WIN = D * (R - 1) + kTaps;  // input samples one thread touches
    __host__ __device__ static constexpr int phys(int i) { return i + i / ROW; }
};

// Geometry of the single-channel kernel.  A thread owns TWO rows of RO consecutive outputs each (rows t and t + NT of
// the tile) and computes them together on the packed fp32x2 pipe form: the tile is staged as PAIRS, entry i =
// (x[i], x[i + NT*ROW]), so that one 128-bit LDS delivers two consecutive samples of both rows as two ready-made
// register pairs, and every tap is one FFMA2 (two for the exact form) with the tap as a broadcast scalar operand.
// Row pitch = 2*ROW floats + a pad that keeps 16-byte alignment and makes pitch/4 odd, so the LDS.128 of a quarter
// warp cover all 32 banks; staging stores are 128-bit too.  RO = 8 for the fused-multiply-add kernels, 4 for the
// exact ones: the unrolled body is then 1208 FFMA2 = 19 KB either way and stays inside the instruction cache (the first
// exact version, 2416 FMUL/FADD = 40 KB, ran with an icc hit rate of 78%: profiles/r1k_kernels.md).
template <int D, int NT, int RO>
struct GeomP {
    static constexpr int TO = 2 * RO * NT;                 // outputs per tile
    static constexpr int ROW = D * RO;                     // input samples between consecutive rows
    static constexpr int HALF = ROW * NT;                  // distance between the two rows of one thread
    static_assert(ROW % 4 == 0 && OFF % 4 == 0, "rows and the tile origin must keep 128-bit alignment");
    static constexpr int PADF = ((2 * ROW) / 4) % 2 == 0 ? 4 : 8;
    static constexpr int PITCHF = 2 * ROW + PADF;          // floats per row of pairs
    static constexpr int PQHI = (D * (RO - 1) + OFF) / 2, PQLO = (OFF - kHist) / 2;  // sample pairs (c, c+1), c = 2*pq, one thread touches
    static constexpr int NP = ((ROW * (NT - 1) + 2 * (PQHI + 1)) + 3) / 4 * 4;       // pair entries staged
    static constexpr int NPQ = NP / 4;
    static constexpr int WORDS = ((NP + ROW - 1) / ROW) * PITCHF;
    __host__ __device__ static constexpr int physf(int i) { return 2 * i + PADF * (i / ROW); }  // float index of entry i
};

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float x, float h) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(x, h)) : fmaf(x, h, acc);
}

// The same tap on a PAIR of lanes with the packed fp32x2 pipe form (FFMA2, sm_100): one instruction, two lanes, the
// tap as a scalar operand broadcast to both.  FFMA2 has the lane throughput of FFMA (bench.py peak kinds 0 and 2), so
// the FP32 pipe time is unchanged, but the pipe is fed with half the issue slots and the other half is free for the
// loads, conversions and constant fetches that otherwise compete with the taps for them.
// EXACT keeps the reference's two roundings: fma(x, h, -0) IS the rounded product (adding -0 changes neither value nor
// the sign of a zero) and fma(p, 1, acc) IS the rounded sum.  The -0 and the 1 arrive as kernel parameters: there is no
// packed FMUL/FADD in SASS, and with literal constants ptxas folds the pair into a single FFMA2 -- one rounding
// (checked with cuobjdump on bench kind 3).
struct Exact2 {
    float negzero, one;
};
template <bool EXACT>
__device__ __forceinline__ float2 mac2(float2 acc, float2 x, float h, const Exact2 &c) {
    if (EXACT) {
        const float2 p = __ffma2_rn(x, make_float2(h, h), make_float2(c.negzero, c.negzero));
        return __ffma2_rn(p, make_float2(c.one, c.one), acc);
    }
    return __ffma2_rn(x, make_float2(h, h), acc);
}

// ------------------------------------------------------------------------------------------------------------------
// input formation
// ------------------------------------------------------------------------------------------------------------------
struct FirDev {
    const float *x, *x2;
    float *y;
    const float *zi;  // points at the 150 live entries of this launch's state: zi_live[s*nzi + (nzi-150) ...]
    float *zi_out;    // the whole state [S][nzi]: rewritten from the last block's input by the one CTA per stream that reads it
    long long ldx, ldy;
    int nzi, n, ny, n_blocks;
    int zi_first;     // first state entry to rewrite (nzi - 150 when only the live entries matter, else 0)
};

template <int KIND>
__device__ __forceinline__ float form(float a, float b, bool in_block) {
    if (KIND == SRC_PLAIN) return a;
    if (KIND == SRC_SQUARE) return __fmul_rn(a, a);
    if (KIND == SRC_MIX_LATE) return __fmul_rn(a, b);
    if (KIND == SRC_PROD_HALF) return in_block ? a : __fmul_rn(a, 0.5f);  // the sum is doubled at the end
    return in_block ? __fmul_rn(__fmul_rn(a, b), 2.0f) : __fmul_rn(a, b);  // SRC_MIX_HALF
}

// value of the FIR input at logical position p (relative to the start of block b) of stream-row xs / x2s
template <int KIND>
__device__ __forceinline__ float source(const FirDev &a, const float *xs, const float *x2s, const float *zs, int b, int p) {
    if (p >= a.n || p < -kHist) return 0.0f;  // outside the filter's support: tile padding only
    long long q;
    bool in_block = p >= 0;
    if (in_block) {
        q = (long long)b * a.n + p;
    } else if (b > 0) {
        q = (long long)b * a.n + p - ((KIND == SRC_MIX_HALF || KIND == SRC_PROD_HALF) ? 0 : 1);  // one-late history, except the mixer (Q8)
    } else {
        const float z = zs[kHist + p];  // carried state already holds the formed value (for the mixer: the product without its x2)
        return KIND == SRC_PROD_HALF ? __fmul_rn(z, 0.5f) : z;
    }
    float v = xs[q];
    float w = (KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF) ? x2s[q] : 0.0f;
    return form<KIND>(v, w, in_block);
}

// The state after the launch, zi[i] = formed(x[N - nzi - 1 + i]) of the LAST block (mixer kinds: N - nzi + i, the plain product):
// written by the CTA that read the old state -- tile 0 of block 0 is the only one whose window reaches into it -- after the barrier
// that ends its staging.  The input is never written by the launch, so no other CTA races with this (a separate 8 us state kernel
// behind each of the seven filters of a step was 1 % of the step).
template <int KIND>
__device__ __forceinline__ void update_state(const FirDev &a, int s, int nthreads) {
    const float *xs = a.x + (long long)s * a.ldx + (long long)(a.n_blocks - 1) * a.n;
    const float *x2s = a.x2 ? a.x2 + (long long)s * a.ldx + (long long)(a.n_blocks - 1) * a.n : nullptr;
    for (int i = a.zi_first + threadIdx.x; i < a.nzi; i += nthreads) {
        const int p = a.n - a.nzi + i - ((KIND == SRC_MIX_HALF || KIND == SRC_PROD_HALF) ? 0 : 1);
        if (p < 0) continue;  // block shorter than the state: the entry keeps its old value (never live for a 151-tap filter)
        const float v = xs[p];
        const float w = (KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF) ? x2s[p] : 0.0f;
        a.zi_out[(long long)s * a.nzi + i] = KIND == SRC_PROD_HALF ? v : form<KIND>(v, w, false);  // the state keeps the plain product
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the register-tiled tap loop shared by every kernel: W(c) returns the staged sample c positions after the start of
// the thread's window (c = D*r - k + OFF for output r, tap k), ACC(r, v, h) accumulates
// ------------------------------------------------------------------------------------------------------------------
#define FMRX_TAP_LOOP(D, LOADV, MACV)                                 \
    _Pragma("unroll") for (int i_ = 0; i_ < D * (R - 1) + kTaps; ++i_) { \
        const int p_ = D * (R - 1) - i_; /* newest sample first */    \
        const int c_ = p_ + OFF;                                      \
        LOADV(c_ + c_ / (D * R));                                     \
        _Pragma("unroll") for (int r = 0; r < R; ++r) {               \
            const int k = D * r - p_;                                 \
            if (k >= 0 && k < kTaps) { MACV(r, k); }                  \
        }                                                             \
    }

// ------------------------------------------------------------------------------------------------------------------
// single-channel kernel
// ------------------------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float4 stage_quad(const FirDev &a, const float *xs, const float *x2s, const float *zs, const float *xt, const float *x2t,
                                             int b, int P0, int i, bool vec_ok) {
    // four consecutive input samples starting at tile index i (block position P0 + i), formed
    const int p = P0 + i;
    if (vec_ok && p >= 0 && p + 4 <= a.n) {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(xt + i));
        float4 w = v;
        if (KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF) w = __ldg(reinterpret_cast<const float4 *>(x2t + i));
        return make_float4(form<KIND>(v.x, w.x, true), form<KIND>(v.y, w.y, true), form<KIND>(v.z, w.z, true), form<KIND>(v.w, w.w, true));
    }
    // outside the block (the 168 history samples of a block's first tile, padding past the end): sample by sample
    return make_float4(source<KIND>(a, xs, x2s, zs, b, p), source<KIND>(a, xs, x2s, zs, b, p + 1), source<KIND>(a, xs, x2s, zs, b, p + 2),
                       source<KIND>(a, xs, x2s, zs, b, p + 3));
}

template <int D, int KIND, bool EXACT, int NT>
__global__ void __launch_bounds__(NT) fir151_kernel(const FirDev a, const __grid_constant__ Taps taps, const Exact2 ex) {
    constexpr int RO = EXACT ? 4 : 8;
    using G = GeomP<D, NT, RO>;
    __shared__ __align__(16) float sm[G::WORDS];
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * G::TO;
    const float *xs = a.x + (long long)s * a.ldx;
    const float *x2s = a.x2 ? a.x2 + (long long)s * a.ldx : nullptr;
    const float *zs = a.zi + (long long)s * a.nzi;
    const int P0 = D * n0 - OFF;  // block position of tile index 0: a multiple of 4, like n, so a quad is wholly inside the block or wholly outside
    constexpr bool MIX = KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF;
    const float *xt = xs + (long long)b * a.n + P0;
    const float *x2t = MIX ? x2s + (long long)b * a.n + P0 : xt;
    const bool vec_ok = (a.n & 3) == 0 && (((uintptr_t)xt | (uintptr_t)x2t) & 15) == 0;
    {
        // entries 4j..4j+3 pair tile samples 4j.. with tile samples HALF+4j..: two 128-bit loads (per input signal), two
        // 128-bit stores.  Unrolled, so that all loads of a thread are in flight before its first store.
        constexpr int ITERS = (G::NPQ + NT - 1) / NT;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int j = threadIdx.x + it * NT;
            if (j >= G::NPQ) break;
            const float4 lo = stage_quad<KIND>(a, xs, x2s, zs, xt, x2t, b, P0, 4 * j, vec_ok);
            const float4 hi = stage_quad<KIND>(a, xs, x2s, zs, xt, x2t, b, P0, 4 * j + G::HALF, vec_ok);
            float4 *dst = reinterpret_cast<float4 *>(sm + G::physf(4 * j));
            dst[0] = make_float4(lo.x, hi.x, lo.y, hi.y);
            dst[1] = make_float4(lo.z, hi.z, lo.w, hi.w);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && b == 0 && a.zi_out) update_state<KIND>(a, s, NT);

    // row t's window starts at entry ROW*t = float PITCHF*t; c = p + OFF is the index inside the window of the sample p
    // positions after the row's first output's newest one, and output r uses it with tap k = D*r - p
    const float *w = sm + G::PITCHF * threadIdx.x;
    float2 acc[RO];  // .x: row t, .y: row t + NT
#pragma unroll
    for (int r = 0; r < RO; ++r) acc[r] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int pq = G::PQHI; pq >= G::PQLO; --pq) {
        const float4 v = *reinterpret_cast<const float4 *>(w + G::physf(2 * pq));
#pragma unroll
        for (int e = 1; e >= 0; --e) {  // newest sample first
            const float2 x2 = e ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
            const int p_ = 2 * pq + e - OFF;
#pragma unroll
            for (int r = 0; r < RO; ++r) {
                const int k = D * r - p_;
                if (k >= 0 && k < kTaps) acc[r] = mac2<EXACT>(acc[r], x2, taps.h[k], ex);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int o = n0 + RO * (threadIdx.x + h * NT);
        float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny + o;
        float out[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r) {
            out[r] = h ? acc[r].y : acc[r].x;
            if (KIND == SRC_PROD_HALF) out[r] = __fmul_rn(out[r], 2.0f);
        }
        if (o + RO <= a.ny && ((reinterpret_cast<uintptr_t>(ys) & 15) == 0)) {
#pragma unroll
            for (int r = 0; r < RO; r += 4) reinterpret_cast<float4 *>(ys)[r / 4] = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
        } else {
#pragma unroll
            for (int r = 0; r < RO; ++r)
                if (o + r < a.ny) ys[r] = out[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Several D = 1 filters of ONE input in one launch (the pilot, 22-54 kHz and 54-60 kHz band-pass filters all read the
// discriminator output): the tile is staged once and the tap loop runs once per filter, the taps of filter f coming from the
// parameter block at a warp-uniform offset, so the loop body exists once in the instruction stream.  Same arithmetic and
// summation order as fir151_kernel<1, SRC_PLAIN, EXACT, NT> -- bit-identical outputs -- minus two of the three stagings.
// All filters carry the same history (the input's tail), so it is staged from filter 0's state; every filter's state is rewritten.
// ------------------------------------------------------------------------------------------------------------------
constexpr int MAXF = 3;
struct TapsF {
    float h[MAXF][kTaps + 1];
};
struct MultiDev {
    const float *x;
    float *y[MAXF];
    float *zi[MAXF];   // [S][nzi] each; the last 150 entries are live
    long long ldx, ldy;
    int nzi, n, ny, n_blocks, nf, zi_first;
};

template <bool EXACT, int NT>
__global__ void __launch_bounds__(NT) fir151_multi_kernel(const MultiDev a, const __grid_constant__ TapsF taps, const Exact2 ex) {
    constexpr int RO = EXACT ? 4 : 8, D = 1;
    using G = GeomP<D, NT, RO>;
    __shared__ __align__(16) float sm[G::WORDS];
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * G::TO;
    FirDev one{};
    one.x = a.x; one.x2 = nullptr; one.zi = a.zi[0] + (a.nzi - kHist); one.ldx = a.ldx; one.nzi = a.nzi; one.n = a.n; one.ny = a.ny; one.n_blocks = a.n_blocks;
    const float *xs = a.x + (long long)s * a.ldx;
    const float *zs = one.zi + (long long)s * a.nzi;
    const int P0 = D * n0 - OFF;
    const float *xt = xs + (long long)b * a.n + P0;
    const bool vec_ok = (a.n & 3) == 0 && ((uintptr_t)xt & 15) == 0;
    {
        constexpr int ITERS = (G::NPQ + NT - 1) / NT;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int j = threadIdx.x + it * NT;
            if (j >= G::NPQ) break;
            const float4 lo = stage_quad<SRC_PLAIN>(one, xs, nullptr, zs, xt, xt, b, P0, 4 * j, vec_ok);
            const float4 hi = stage_quad<SRC_PLAIN>(one, xs, nullptr, zs, xt, xt, b, P0, 4 * j + G::HALF, vec_ok);
            float4 *dst = reinterpret_cast<float4 *>(sm + G::physf(4 * j));
            dst[0] = make_float4(lo.x, hi.x, lo.y, hi.y);
            dst[1] = make_float4(lo.z, hi.z, lo.w, hi.w);
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && b == 0) {
        const float *xl = xs + (long long)(a.n_blocks - 1) * a.n;
        for (int i = a.zi_first + threadIdx.x; i < a.nzi; i += NT) {
            const int p = a.n - a.nzi + i - 1;
            if (p < 0) continue;
            const float v = xl[p];
            for (int f = 0; f < a.nf; ++f) a.zi[f][(long long)s * a.nzi + i] = v;
        }
    }
    const float *w = sm + G::PITCHF * threadIdx.x;
#pragma unroll 1
    for (int f = 0; f < a.nf; ++f) {
        const float *h = taps.h[f];
        float2 acc[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r) acc[r] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int pq = G::PQHI; pq >= G::PQLO; --pq) {
            const float4 v = *reinterpret_cast<const float4 *>(w + G::physf(2 * pq));
#pragma unroll
            for (int e = 1; e >= 0; --e) {
                const float2 x2 = e ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
                const int p_ = 2 * pq + e - OFF;
#pragma unroll
                for (int r = 0; r < RO; ++r) {
                    const int k = D * r - p_;
                    if (k >= 0 && k < kTaps) acc[r] = mac2<EXACT>(acc[r], x2, h[k], ex);
                }
            }
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int o = n0 + RO * (threadIdx.x + hh * NT);
            float *ys = a.y[f] + (long long)s * a.ldy + (long long)b * a.ny + o;
            float out[RO];
#pragma unroll
            for (int r = 0; r < RO; ++r) out[r] = hh ? acc[r].y : acc[r].x;
            if (o + RO <= a.ny && ((reinterpret_cast<uintptr_t>(ys) & 15) == 0)) {
#pragma unroll
                for (int r = 0; r < RO; r += 4) reinterpret_cast<float4 *>(ys)[r / 4] = make_float4(out[r], out[r + 1], out[r + 2], out[r + 3]);
            } else {
#pragma unroll
                for (int r = 0; r < RO; ++r)
                    if (o + r < a.ny) ys[r] = out[r];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// pllCombine's filter with the reference's own arithmetic (src/helper.cpp:139): an in-block tap is
//     y[n] += pow(x[p], 2) * h[k]     -- the square and the product are DOUBLE, y[n] is a float,
// i.e. acc <- (float)((double)acc + ((double)x * (double)x) * (double)h[k]); a history tap (:144) is the plain
// float multiply-add on the stored float square.  Every partial sum is therefore rounded to 53 bits and then to 24.
//
// Per tap: DMUL, DADD, and the rounding to the nearest float (ties to even) done without a conversion instruction: the
// accumulator stays a double that holds a float value, and RN24(s) = (s + M) - M with M = 1.5 * 2^(e + 29), e the
// exponent of s -- the first DADD rounds s on the grid ulp53(M) = 2^(e - 23), which is the float grid at s (M's own
// mantissa is even on that grid, so ties go to even exactly as in the conversion), the second DADD is exact.  Clamping
// e at -126 makes the grid 2^-149 below the smallest normal float, which is the conversion's gradual underflow.  M is
// built from the high word of s with three integer instructions (first version: the whole rounding on the bit pattern,
// 6-7 integer instructions per tap on the half-rate ALU pipe: 3.7 ms per step against 2 ms of FP64 pipe).
// Two things the identity does not cover are excluded by bounds checked per tile and on the taps -- x = 0 or 1e-8 <= |x| <= 1e12,
// h = 0 or 1e-12 <= |h| <= 1e10: sums cannot overflow a float (151 * 1e34 < FLT_MAX), and a non-zero sum cannot round to
// zero, whose sign the identity would lose ((s + M) - M = +0 for a tiny negative s where the conversion gives -0): every
// non-zero product is >= 2^-93, so every partial sum is a multiple of 2^-149 and representable down to the last subnormal.
// A tile or tap set outside the bounds, and a thread whose window reaches into the history (first tile of a block), take
// the expression as written, tap by tap, with the conversion instructions.
// Tile = 128 threads x 8 consecutive outputs; squares staged once per CTA as doubles (pitch 9: conflict-free LDS.64).
// ------------------------------------------------------------------------------------------------------------------
struct TapsD {
    double h[kTaps + 1];
};

constexpr int SQ_NT = 128, SQ_RO = 8, SQ_LEAD = 152;           // staged samples ahead of the tile's first output (>= 150, multiple of 4)
constexpr int SQ_N = SQ_NT * SQ_RO + SQ_LEAD;                  // staged samples per tile
__host__ __device__ constexpr int sq_phys(int i) { return i + i / SQ_RO; }

__device__ __forceinline__ double round_to_float_kept_double(double s) {
    const int e = max(__double2hiint(s) & 0x7FF00000, 0x38100000);       // exponent field, not below 2^-126
    const double M = __hiloint2double(e + 0x01D80000, 0);                // 1.5 * 2^(e + 29)
    return __dsub_rn(__dadd_rn(s, M), M);
}

__global__ void __launch_bounds__(SQ_NT) fir151_sq_exact_kernel(const FirDev a, const __grid_constant__ Taps taps, const __grid_constant__ TapsD tapsd, int taps_bounded) {
    __shared__ double q[sq_phys(SQ_N) + 1];    // in-block: x^2 (exact in double); history: the stored float square
    __shared__ double hd[kTaps + 1];
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * (SQ_NT * SQ_RO);
    const float *xs = a.x + (long long)s * a.ldx;
    const float *zs = a.zi + (long long)s * a.nzi;
    const int P0 = n0 - SQ_LEAD;
    bool finite = true;
    for (int i = threadIdx.x; i < SQ_N; i += SQ_NT) {
        const int p = P0 + i;
        double v = 0.0;
        if (p >= 0 && p < a.n) {
            const double x = (double)xs[(long long)b * a.n + p];
            v = __dmul_rn(x, x);
            finite = finite && (v == 0.0 || (v >= 1.0e-16 && v <= 1.0e24));  // false for inf and NaN too
        } else if (p < 0 && p >= -kHist) {
            v = (double)source<SRC_SQUARE>(a, xs, nullptr, zs, b, p);
        }
        q[sq_phys(i)] = v;
    }
    for (int i = threadIdx.x; i <= kTaps; i += SQ_NT) hd[i] = tapsd.h[i];
    const bool tile_finite = __syncthreads_and(finite);
    if (blockIdx.x == 0 && b == 0 && a.zi_out) update_state<SRC_SQUARE>(a, s, SQ_NT);

    const int t = threadIdx.x, o0 = n0 + SQ_RO * t;   // first output of this thread
    float out[SQ_RO];
    const bool slow = o0 < kHist || !tile_finite || !taps_bounded;
    if (!slow) {
        // sample j of the walk is block position o0 + 7 - j = staged index 8t + 159 - j; output r takes it with tap r - 7 + j
        const int base = SQ_RO * t + SQ_LEAD + SQ_RO - 1;
        double acc[SQ_RO];
#pragma unroll
        for (int r = 0; r < SQ_RO; ++r) acc[r] = 0.0;
        // ramp up: outputs join one by one (static tap validity)
#pragma unroll
        for (int j = 0; j < SQ_RO; ++j) {
            const double x2 = q[sq_phys(base - j)];
#pragma unroll
            for (int r = SQ_RO - 1 - j; r < SQ_RO; ++r) {
                acc[r] = round_to_float_kept_double(__dadd_rn(acc[r], __dmul_rn(x2, hd[r - (SQ_RO - 1) + j])));
            }
        }
        // steady state: all eight outputs take every sample; the eight taps in flight slide by one per sample
        double hw[SQ_RO];
#pragma unroll
        for (int r = 0; r < SQ_RO; ++r) hw[r] = hd[r + 1];     // taps of j = 8: r - 7 + 8
#pragma unroll 8
        for (int j = SQ_RO; j < kTaps; ++j) {
            const double x2 = q[sq_phys(base - j)];
#pragma unroll
            for (int r = 0; r < SQ_RO; ++r) acc[r] = round_to_float_kept_double(__dadd_rn(acc[r], __dmul_rn(x2, hw[r])));
#pragma unroll
            for (int r = 0; r < SQ_RO - 1; ++r) hw[r] = hw[r + 1];
            hw[SQ_RO - 1] = hd[j + 1];                             // hd[151] = 0: loaded after the last steady sample, never used
        }
        // ramp down: outputs leave one by one
#pragma unroll
        for (int j = kTaps; j < kTaps + SQ_RO - 1; ++j) {
            const double x2 = q[sq_phys(base - j)];
#pragma unroll
            for (int r = 0; r < SQ_RO - 1 - (j - kTaps); ++r)
                acc[r] = round_to_float_kept_double(__dadd_rn(acc[r], __dmul_rn(x2, hd[r - (SQ_RO - 1) + j])));
        }
#pragma unroll
        for (int r = 0; r < SQ_RO; ++r) out[r] = (float)acc[r];
    }
    if (slow) {
        // the expression as written, with conversion instructions: history taps in float, in-block taps through double
        for (int r = 0; r < SQ_RO; ++r) {
            const int n = o0 + r;
            float acc = 0.0f;
            for (int k = 0; k < kTaps; ++k) {
                const int p = n - k;
                const double v = q[sq_phys(p - P0)];
                if (p >= 0) acc = (float)__dadd_rn((double)acc, __dmul_rn(v, hd[k]));
                else acc = __fadd_rn(acc, __fmul_rn((float)v, taps.h[k]));
            }
            out[r] = acc;
        }
    }
    float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny + o0;
    if (o0 + SQ_RO <= a.ny && ((reinterpret_cast<uintptr_t>(ys) & 15) == 0)) {
        reinterpret_cast<float4 *>(ys)[0] = make_float4(out[0], out[1], out[2], out[3]);
        reinterpret_cast<float4 *>(ys)[1] = make_float4(out[4], out[5], out[6], out[7]);
    } else {
#pragma unroll
        for (int r = 0; r < SQ_RO; ++r)
            if (o0 + r < a.ny) ys[r] = out[r];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// two-channel kernel (I and Q share the taps): float inputs, or raw u8 with the discriminator fused behind it
// ------------------------------------------------------------------------------------------------------------------
struct IqDev {
    const float *xi, *xq;     // float inputs (fir_iq) ...
    const uint8_t *raw;       // ... or interleaved u8 (front end)
    float *yi, *yq, *demod;
    const float *zii, *ziq;   // [S][150]
    long long ldx, ldy;       // ldx in complex samples for floats, in BYTES for raw
    int n, ny, n_blocks;
};

__device__ __forceinline__ float u8_to_f32(unsigned v) {
    // (v - 128) / 128 exactly: 0x4B000000|v is 2^23 + v; one FMA scales and removes the bias with a single (exact) rounding
    return fmaf(__uint_as_float(0x4B000000u | v), 0.0078125f, -65537.0f);
}

// byte K of a packed word -> float: one PRMT builds 0x4B0000bb, one FMA finishes (same exact result as u8_to_f32)
template <int K>
__device__ __forceinline__ float u8_lane_to_f32(unsigned packed) {
    return fmaf(__uint_as_float(__byte_perm(packed, 0x4B000000u, 0x7440 | K)), 0.0078125f, -65537.0f);
}

template <bool RAW>
__device__ __forceinline__ float2 source_iq(const IqDev &a, int s, int b, int p) {
    if (p >= a.n || p < -kHist) return make_float2(0.0f, 0.0f);  // outside the filter's support: tile padding only
    if (p < 0 && b == 0) return make_float2(a.zii[(long long)s * kHist + kHist + p], a.ziq[(long long)s * kHist + kHist + p]);
    const long long q = (long long)b * a.n + p - (p < 0 ? 1 : 0);
    if (RAW) {
        const uchar2 v = *reinterpret_cast<const uchar2 *>(a.raw + (long long)s * a.ldx + 2 * q);
        return make_float2(u8_to_f32(v.x), u8_to_f32(v.y));
    }
    return make_float2(a.xi[(long long)s * a.ldx + q], a.xq[(long long)s * a.ldx + q]);
}

// src/rf_module.cpp:19-33 with its operand types: fp32 numerator, double denominator and quotient
__device__ __forceinline__ float discriminate(float i, float q, float pi_, float pq_) {
    const double den = __dadd_rn(__dmul_rn((double)i, (double)i), __dmul_rn((double)q, (double)q));
    if (den == 0.0) return 0.0f;
    const float num = __fsub_rn(__fmul_rn(i, __fsub_rn(q, pq_)), __fmul_rn(q, __fsub_rn(i, pi_)));
    return (float)__ddiv_rn((double)num, den);
}

template <int D, bool RAW, bool EXACT>
__global__ void __launch_bounds__(CTQ) fir151_iq_kernel(const IqDev a, const __grid_constant__ Taps taps) {
    using G = Geom<D, CTQ>;
    __shared__ float2 smq[G::WORDS];
    __shared__ float2 edge[CTQ + 1];  // edge[t+1] = last output of thread t
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * G::TO;
    const int P0 = D * n0 - OFF;
    constexpr int OCTS = (G::SPAN + 7) / 8;
    const uint8_t *rt = RAW ? a.raw + (long long)s * a.ldx + 2LL * b * a.n + 2LL * P0 : nullptr;
    if (RAW && P0 >= 0 && P0 + 8 * OCTS <= a.n && ((uintptr_t)rt & 15) == 0) {
        // interior tile: eight complex samples (16 bytes) per lane per load; ROW is a multiple of 8, so the eight land in
        // one padded row and share the row offset
        static_assert(G::ROW % 8 == 0, "an aligned group of 8 samples must not straddle a padded row");
        for (int j = threadIdx.x; j < OCTS; j += CTQ) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rt) + j);
            const int i = 8 * j;
            float2 *dst = smq + G::phys(i);
            dst[0] = make_float2(u8_lane_to_f32<0>(v.x), u8_lane_to_f32<1>(v.x));
            if (i + 1 < G::SPAN) dst[1] = make_float2(u8_lane_to_f32<2>(v.x), u8_lane_to_f32<3>(v.x));
            if (i + 2 < G::SPAN) dst[2] = make_float2(u8_lane_to_f32<0>(v.y), u8_lane_to_f32<1>(v.y));
            if (i + 3 < G::SPAN) dst[3] = make_float2(u8_lane_to_f32<2>(v.y), u8_lane_to_f32<3>(v.y));
            if (i + 4 < G::SPAN) dst[4] = make_float2(u8_lane_to_f32<0>(v.z), u8_lane_to_f32<1>(v.z));
            if (i + 5 < G::SPAN) dst[5] = make_float2(u8_lane_to_f32<2>(v.z), u8_lane_to_f32<3>(v.z));
            if (i + 6 < G::SPAN) dst[6] = make_float2(u8_lane_to_f32<0>(v.w), u8_lane_to_f32<1>(v.w));
            if (i + 7 < G::SPAN) dst[7] = make_float2(u8_lane_to_f32<2>(v.w), u8_lane_to_f32<3>(v.w));
        }
    } else {
        for (int i = threadIdx.x; i < G::SPAN; i += CTQ) smq[G::phys(i)] = source_iq<RAW>(a, s, b, P0 + i);
    }
    __syncthreads();

    const float2 *w = smq + G::PITCH * threadIdx.x;
    float ai[R], aq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) ai[r] = aq[r] = 0.0f;
    float2 v;
#define LOADV(idx) v = w[idx]
#define MACV(r, k) ai[r] = mac<EXACT>(ai[r], v.x, taps.h[k]); aq[r] = mac<EXACT>(aq[r], v.y, taps.h[k])
    FMRX_TAP_LOOP(D, LOADV, MACV)
#undef LOADV
#undef MACV

    const int o = n0 + R * threadIdx.x;
    const long long base = (long long)s * a.ldy + (long long)b * a.ny + o;
    if (a.yi) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (o + r < a.ny) { a.yi[base + r] = ai[r]; a.yq[base + r] = aq[r]; }
    }
    if (RAW) {
        // discriminator: the sample before this thread's first output is the neighbour's last one.  Thread 0 takes
        // zero: right at a block start (Q3); for every other tile demod[n0] is rewritten by demod_edge_kernel.
        edge[threadIdx.x + 1] = make_float2(ai[R - 1], aq[R - 1]);
        if (threadIdx.x == 0) edge[0] = make_float2(0.0f, 0.0f);
        __syncthreads();
        float2 prev = edge[threadIdx.x];
        float d[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            d[r] = discriminate(ai[r], aq[r], prev.x, prev.y);
            prev = make_float2(ai[r], aq[r]);
        }
        float *ys = a.demod + base;
        if (o + R <= a.ny && ((reinterpret_cast<uintptr_t>(ys) & 15) == 0)) {
            reinterpret_cast<float4 *>(ys)[0] = make_float4(d[0], d[1], d[2], d[3]);
            reinterpret_cast<float4 *>(ys)[1] = make_float4(d[4], d[5], d[6], d[7]);
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (o + r < a.ny) ys[r] = d[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// RF front end: u8 IQ -> 151-tap LPF /10 on I and Q -> discriminator, as a register-streaming FIR.
//
// One THREAD owns a run of L consecutive outputs of one (stream, block) and walks its input newest-first in rows of
// ten complex samples (20 bytes).  An output is "open" while the walk is inside its 151-sample support, i.e. for 16
// rows, so 16 (+1 while two rows are in flight) complex accumulators live in registers; per row every tap is used
// exactly once (151 MACs on I, 151 on Q), each input byte is converted once, and nothing goes through shared memory.
// For one output the walk visits its samples newest-first = taps in ascending order, the reference's summation order.
// The loop body covers two rows (~1.4k instructions, 22 KB: resident in the 32 KB L1.5 instruction cache, which the
// fully unrolled tile kernel -- 85 KB -- was not; ncu: stalled_no_instruction 3.5 per issue, ICC hit rate 52%).
// Rows are offset by one sample (row rho = samples 10*rho-8 .. 10*rho+1) so that four rows = 80 bytes start on a
// 16-byte boundary: input arrives as five 128-bit loads per four rows, issued half a group ahead of their use.
// The run is extended by one output below (for the discriminator's previous sample) and by the 15 rows in which the
// lowest outputs finish: L+16 rows for L outputs (L = 240: 94 % useful).
// ------------------------------------------------------------------------------------------------------------------
constexpr int SLOTS = 17;

// taps regrouped for the row walk: t[j][a] = h[10*a - 1 + j] (0 where that index is outside 0..150), so that the 16 taps
// sample j of a row needs are contiguous and reach the uniform registers as 128-bit constant loads
struct RowTaps {
    float t[10][16];
};

template <bool EXACT>
__global__ void __launch_bounds__(64, 8) frontend_stream_kernel(const IqDev a, const __grid_constant__ RowTaps taps, int L, int segs, long long total, const Exact2 ex) {
    const long long gid = blockIdx.x * 64LL + threadIdx.x;
    if (gid >= total) return;
    const int g = (int)(gid % segs);
    const long long sb = gid / segs;
    const int b = (int)(sb % a.n_blocks), s = (int)(sb / a.n_blocks);
    const int a0 = g * L;                                   // first output of the run (multiple of 4)
    const int hi = min(a0 + L, a.ny);                       // one past the last output stored
    const uint8_t *row = a.raw + (long long)s * a.ldx + 2LL * b * a.n;
    const bool aligned = ((uintptr_t)row & 15) == 0;
    const long long ob = (long long)s * a.ldy + (long long)b * a.ny;

    float2 bb[SLOTS];  // (I, Q) accumulators of the open outputs: one packed lane pair each
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) bb[i] = make_float2(0.0f, 0.0f);
    float2 ycur = make_float2(0.0f, 0.0f);                  // the output completed just before (index n+1 when n completes)

    // group of four rows with top row rt (rt = 3 mod 4): samples 10*rt-38 .. 10*rt+1, bytes 20*rt-76 .. 20*rt+4
    uint4 cur[5];
    bool fast_cur;
    auto load_group = [&](int rt) {
        const int lo = 10 * rt - 38;
        fast_cur = aligned && lo >= 0 && lo + 40 <= a.n;
        if (fast_cur) {
            const uint4 *p = reinterpret_cast<const uint4 *>(row + 2LL * lo);
#pragma unroll
            for (int i = 0; i < 5; ++i) cur[i] = __ldg(p + i);
        }
    };
    int rho = a0 + L - 1;                                   // top row of the run
    load_group(rho);
    const int iters = (L + 16) / 2;
    for (int it = 0; it < iters; ++it, rho -= 2) {
        // this iteration: rows rho (upper) and rho-1 = the upper (even it) or lower (odd it) ten words of the group
        const bool lower = it & 1;
        const unsigned *cw = reinterpret_cast<const unsigned *>(cur);
        unsigned w[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) w[i] = lower ? cw[i] : cw[10 + i];
        const bool fast = fast_cur;
        if (lower && it + 1 < iters) load_group(rho - 2);   // the group is consumed: fetch the next one under this iteration's math
        float2 yh, yl;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            float2 smp[10];                                 // smp[j] = sample 10*(rho-rr) + 1 - j
            if (fast) {
#pragma unroll
                for (int q = 0; q < 5; ++q) {               // word q of this row holds samples (base + 2q, base + 2q + 1)
                    const unsigned u = w[5 * (1 - rr) + q];
                    smp[9 - 2 * q] = make_float2(u8_lane_to_f32<0>(u), u8_lane_to_f32<1>(u));
                    smp[8 - 2 * q] = make_float2(u8_lane_to_f32<2>(u), u8_lane_to_f32<3>(u));
                }
            } else {  // history, block edges, unaligned rows: rare, one sample at a time
#pragma unroll 1
                for (int q = 0; q < 10; ++q) {
                    const float2 v = source_iq<true>(a, s, b, 10 * (rho - rr) + 1 - q);
#pragma unroll
                    for (int z = 0; z < 10; ++z)
                        if (z == q) smp[z] = v;
                }
            }
#pragma unroll
            for (int j = 0; j < 10; ++j) {
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) {
                    const int k = 10 * sl - 1 + j;
                    if (k >= 0 && k < kTaps) {
                        const int i = sl + 1 - rr;
                        bb[i] = mac2<EXACT>(bb[i], smp[j], taps.t[j][sl], ex);
                    }
                }
            }
            if (rr == 0) yh = bb[16];  // output rho+15
            else yl = bb[15];          // output rho+14
        }
#pragma unroll
        for (int i = SLOTS - 1; i >= 2; --i) bb[i] = bb[i - 2];
        bb[0] = bb[1] = make_float2(0.0f, 0.0f);

        // outputs rho+15 (yh) and rho+14 (yl) are complete; demod[n] pairs output n with output n-1.  An output is
        // genuine only if its first row lay inside the walk, i.e. its index is below a0+L.
        const int ne = rho + 15;                            // even
        if (a.yi) {
            if (ne >= a0 && ne < hi) { a.yi[ob + ne] = yh.x; a.yq[ob + ne] = yh.y; }
            if (ne - 1 >= a0 && ne - 1 < hi) { a.yi[ob + ne - 1] = yl.x; a.yq[ob + ne - 1] = yl.y; }
        }
        if (ne + 2 <= a0 + L && ne >= a0) {                 // demod[ne+1] (needs ycur = y[ne+1]) and demod[ne]
            const float d1 = discriminate(ycur.x, ycur.y, yh.x, yh.y);
            const float2 pv = ne == 0 ? make_float2(0.0f, 0.0f) : yl;  // block start: previous sample is zero (Q3)
            const float d0 = discriminate(yh.x, yh.y, pv.x, pv.y);
            float *yd = a.demod + ob + ne;
            if (ne + 1 < hi && ((uintptr_t)yd & 7) == 0) *reinterpret_cast<float2 *>(yd) = make_float2(d0, d1);
            else { if (ne < hi) yd[0] = d0; if (ne + 1 < hi) yd[1] = d1; }
        }
        ycur = yl;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Second form of the streaming front end (the default; the kernel above is kept for unaligned / ragged inputs and as
// FMRX_FRONTEND=1).  ncu on the first form (profiles/r2j_kernels.md): 0.6 non-FFMA2 instructions per FFMA2 -- ptxas
// fetched all 151 taps at the top of every two-row iteration (78 uniform + 60 vector registers), spilled ~28 live
// values around the tap loop to make room, kept the input group in local memory, and carried a second copy of the loop
// (the per-sample path for history / edge rows) whose call sites pin more registers.  This form removes all three:
//   * taps live in shared memory and are fetched per ROW with explicit 128-bit loads (40 LDS.128 per 302 FFMA2,
//     broadcast, conflict-free), 16 registers at a time.  The reference's summation order forces rows outer / samples
//     inner, so a tap cannot be reused across rows without holding all 151; reloading is the cheap side of that trade.
//   * one iteration walks a whole 80-byte group (four rows): no half-group selects, the five 128-bit loads of the next
//     group refill the registers of the current one as soon as a row has consumed them (>= 300 FFMA2 ahead of use).
//   * no history path: the hot kernel treats everything before the block as silence and does not emit outputs 0..15 of
//     a block (the only ones whose support reaches the history); frontend_edge_kernel computes those 16 from scratch
//     (in-block samples and history, reference order) -- 0.1 % of the MACs.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SLOTS4 = 19;
constexpr int EDGE_OUT = 16;  // outputs per block left to the edge kernel: n < 15 touch the history, 15 keeps quads whole

__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <bool EXACT>
__global__ void __launch_bounds__(64, 8) frontend_stream4_kernel(const IqDev a, const __grid_constant__ RowTaps taps, int L, int segs, long long total, const Exact2 ex) {
    __shared__ __align__(16) float staps[4 * 160];          // one copy per row of the group: identical loads from one address would be merged
    for (int i = threadIdx.x; i < 4 * 160; i += 64) staps[i] = taps.t[(i % 160) / 16][i % 16];
    __syncthreads();
    const long long gid = blockIdx.x * 64LL + threadIdx.x;
    if (gid >= total) return;
    const int g = (int)(gid % segs);
    const long long sb = gid / segs;
    const int b = (int)(sb % a.n_blocks), s = (int)(sb / a.n_blocks);
    const int a0 = g * L;                                   // first output of the run (multiple of 4)
    const int lo_out = a0 == 0 ? EDGE_OUT : a0;             // first output this thread emits
    const uint4 *row = reinterpret_cast<const uint4 *>(a.raw + (long long)s * a.ldx + 2LL * b * a.n);
    const long long ob = (long long)s * a.ldy + (long long)b * a.ny;
    const unsigned tap_base = (unsigned)__cvta_generic_to_shared(staps);

    float2 bb[SLOTS4];
#pragma unroll
    for (int i = 0; i < SLOTS4; ++i) bb[i] = make_float2(0.0f, 0.0f);
    float2 ycur = make_float2(0.0f, 0.0f);                  // y[rho + 16]: the lowest output of the previous iteration

    // group with top row rho (= 3 mod 4): samples 10*rho-38 .. 10*rho+1 = bytes 20*rho-76 .. 20*rho+3, five uint4;
    // quad i is word 4i..4i+3, word w = samples (lo + 2w, lo + 2w + 1) as I,Q,I,Q bytes
    uint4 cur[5];
    const uint4 silence = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);  // u8 128 = exactly 0.0
    auto fetch = [&](int rt, int i) -> uint4 {
        const int q = (5 * rt - 19) / 4 + i;                // (20*rt - 76) / 16, exact: rt = 3 mod 4
        return q >= 0 ? __ldg(row + q) : silence;           // below the block: only the first run gets there, and nothing it emits depends on it
    };
    int rho = a0 + L - 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) cur[i] = fetch(rho, i);
    const int iters = (L + 16) / 4;
    for (int it = 0; it < iters; ++it, rho -= 4) {
        const bool more = it + 1 < iters;
        float2 yo[4];                                       // yo[rr] = y[rho + 15 - rr]
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            unsigned tb = tap_base + 640 * rr;
            asm volatile("" : "+r"(tb));                    // the taps are re-read per row: do not let them be kept (151 registers) or hoisted
            const unsigned *cw = reinterpret_cast<const unsigned *>(cur);
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const int pos = 39 - 10 * rr - j;           // sample index inside the group
                const unsigned u = cw[pos >> 1];
                const float2 rawf = (pos & 1) ? make_float2(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7442)), __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7443)))
                                              : make_float2(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7440)), __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7441)));
                const float2 smp = __ffma2_rn(rawf, make_float2(0.0078125f, 0.0078125f), make_float2(-65537.0f, -65537.0f));  // (v - 128) / 128, exact
                float t[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = lds_f4(tb + (64 * j + 16 * q));
                    t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int sl = 0; sl < 16; ++sl) {
                    const int k = 10 * sl - 1 + j;
                    if (k >= 0 && k < kTaps) bb[sl + 3 - rr] = mac2<EXACT>(bb[sl + 3 - rr], smp, t[sl], ex);
                }
            }
            yo[rr] = bb[18 - rr];
            // this row's quads are consumed: refill them with the next group
            if (more) {
                if (rr == 0) cur[4] = fetch(rho - 4, 4);
                if (rr == 1) cur[3] = fetch(rho - 4, 3);
                if (rr == 2) cur[2] = fetch(rho - 4, 2);
                if (rr == 3) { cur[1] = fetch(rho - 4, 1); cur[0] = fetch(rho - 4, 0); }
            }
        }
#pragma unroll
        for (int i = SLOTS4 - 1; i >= 4; --i) bb[i] = bb[i - 4];
        bb[0] = bb[1] = bb[2] = bb[3] = make_float2(0.0f, 0.0f);

        // outputs rho+12 .. rho+15 are complete; the aligned quad rho+13 .. rho+16 (ycur = y[rho+16]) can be emitted:
        // demod[n] pairs y[n] with y[n-1].  Genuine only inside [lo_out, a0 + L).
        const int q0 = rho + 13;                            // multiple of 4
        if (q0 >= lo_out && q0 + 4 <= a0 + L) {
            const float d3 = discriminate(ycur.x, ycur.y, yo[0].x, yo[0].y);
            const float d2 = discriminate(yo[0].x, yo[0].y, yo[1].x, yo[1].y);
            const float d1 = discriminate(yo[1].x, yo[1].y, yo[2].x, yo[2].y);
            const float d0 = discriminate(yo[2].x, yo[2].y, yo[3].x, yo[3].y);
            *reinterpret_cast<float4 *>(a.demod + ob + q0) = make_float4(d0, d1, d2, d3);
            if (a.yi) {
                *reinterpret_cast<float4 *>(a.yi + ob + q0) = make_float4(yo[2].x, yo[1].x, yo[0].x, ycur.x);
                *reinterpret_cast<float4 *>(a.yq + ob + q0) = make_float4(yo[2].y, yo[1].y, yo[0].y, ycur.y);
            }
        }
        ycur = yo[3];
    }
}

// outputs 0 .. EDGE_OUT-1 of every (stream, block), from scratch in the reference's order (taps ascending: in-block
// samples first, then the history), and their discriminator values; 16 lanes per (stream, block).
template <bool EXACT>
__global__ void __launch_bounds__(128) frontend_edge_kernel(const IqDev a, const __grid_constant__ Taps taps, int n_sb) {
    const int sbi = blockIdx.x * 8 + threadIdx.x / EDGE_OUT, n = threadIdx.x % EDGE_OUT;
    const bool live = sbi < n_sb;
    const int b = live ? sbi % a.n_blocks : 0, s = live ? sbi / a.n_blocks : 0;
    float yi = 0.0f, yq = 0.0f;
    if (live) {
        // the loads do not depend on the sums: fetch eight samples at a time, then accumulate in order
        for (int k0 = 0; k0 < kTaps; k0 += 8) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = k0 + u < kTaps ? source_iq<true>(a, s, b, 10 * n - k0 - u) : make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < kTaps) { yi = mac<EXACT>(yi, v[u].x, taps.h[k0 + u]); yq = mac<EXACT>(yq, v[u].y, taps.h[k0 + u]); }
        }
    }
    float pi_ = __shfl_up_sync(0xffffffffu, yi, 1, EDGE_OUT), pq_ = __shfl_up_sync(0xffffffffu, yq, 1, EDGE_OUT);
    if (n == 0) pi_ = pq_ = 0.0f;                           // block start: the previous sample is zero (Q3)
    if (!live || n >= a.ny) return;
    const long long o = (long long)s * a.ldy + (long long)b * a.ny + n;
    a.demod[o] = discriminate(yi, yq, pi_, pq_);
    if (a.yi) { a.yi[o] = yi; a.yq[o] = yq; }
}

template <bool RAW>
__global__ void iq_state_kernel(const IqDev a) {
    const int s = blockIdx.x, i = threadIdx.x;
    if (i >= kHist) return;
    const float2 v = source_iq<RAW>(a, s, a.n_blocks - 1, a.n - kHist - 1 + i);
    const_cast<float *>(a.zii)[(long long)s * kHist + i] = v.x;
    const_cast<float *>(a.ziq)[(long long)s * kHist + i] = v.y;
}

// ------------------------------------------------------------------------------------------------------------------
// small element-wise kernels: unpack (a1), stand-alone discriminator (a7)
// ------------------------------------------------------------------------------------------------------------------
__global__ void unpack_kernel(const uint8_t *raw, size_t n, float *out) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = u8_to_f32(raw[i]);
}

__global__ void demod_kernel(const float *I, const float *Q, float *out, int n, long long total) {
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= total) return;
    const bool first = (g % n) == 0;
    out[g] = discriminate(I[g], Q[g], first ? 0.0f : I[g - 1], first ? 0.0f : Q[g - 1]);
}

Taps make_taps(const float *h) {
    Taps t;
    for (int k = 0; k < kTaps; ++k) t.h[k] = h[k];
    t.h[kTaps] = 0.0f;
    return t;
}

// threads per CTA: a tile is 2 rows x RO outputs x NT threads = 1024 outputs (exact: RO 4, NT 128; fma: RO 8, NT 64);
// for the 3648-sample RDS blocks 912-output tiles (NT 114 / 57) divide the block exactly
template <bool EXACT> struct TileThreads { static constexpr int STD = EXACT ? 128 : 64, RDS = EXACT ? 114 : 57; };

template <int D, int KIND, bool EXACT>
int launch_fir_dke(const FirJob &j, const FirDev &d, fmrx_stream_t st) {
    const Taps t = make_taps(j.h);
    const Exact2 ex{-0.0f, 1.0f};
    constexpr int RO = EXACT ? 4 : 8;
    using TT = TileThreads<EXACT>;
    if (D == 1 && KIND == SRC_PLAIN && d.ny % (2 * RO * TT::RDS) == 0 && d.ny % (2 * RO * TT::STD) != 0) {
        constexpr int DD = D == 1 && KIND == SRC_PLAIN ? D : 1, KK = D == 1 && KIND == SRC_PLAIN ? KIND : SRC_PLAIN;  // instantiated once only
        dim3 grid(d.ny / (2 * RO * TT::RDS), j.n_blocks, j.n_streams);
        fir151_kernel<DD, KK, EXACT, TT::RDS><<<grid, TT::RDS, 0, st>>>(d, t, ex);
        return (int)cudaGetLastError();
    }
    constexpr int TO = 2 * RO * TT::STD;
    dim3 grid((d.ny + TO - 1) / TO, j.n_blocks, j.n_streams);
    fir151_kernel<D, KIND, EXACT, TT::STD><<<grid, TT::STD, 0, st>>>(d, t, ex);
    return (int)cudaGetLastError();
}

template <int D, int KIND>
int launch_fir_dk(const FirJob &j, const FirDev &d, dim3, fmrx_stream_t st) {
    // D = 5: the exact kernel's tile (4 outputs per row, 128 threads) is also the faster one -- 0.146 ms against 0.357 ms for the mono
    // low-pass of 4096 stations with the fused-multiply-add form (8 outputs per row: a 190-sample window per thread) -- so the
    // decimating low-pass filters keep the reference's rounding whatever the numerics setting
    if constexpr (D == 5) return launch_fir_dke<D, KIND, true>(j, d, st);
    else return j.exact ? launch_fir_dke<D, KIND, true>(j, d, st) : launch_fir_dke<D, KIND, false>(j, d, st);
}

template <int KIND>
int launch_fir_k(const FirJob &j, const FirDev &d, dim3 grid, fmrx_stream_t st) {
    int e;
    // only the (kind, decimation) pairs the receive chain and the function-level API use are instantiated
    if (KIND == SRC_SQUARE && j.exact) {  // reference-exact pllCombine filter: double products rounded into a float sum
        if (j.decim != 1) return (int)cudaErrorInvalidValue;
        TapsD td;
        for (int k = 0; k < kTaps; ++k) td.h[k] = (double)j.h[k];
        td.h[kTaps] = 0.0;
        dim3 g((d.ny + SQ_NT * SQ_RO - 1) / (SQ_NT * SQ_RO), j.n_blocks, j.n_streams);
        int bounded = 1;
        for (int k = 0; k < kTaps; ++k) {
            const float m = j.h[k] < 0 ? -j.h[k] : j.h[k];
            bounded &= (m == 0.0f || (m >= 1e-12f && m <= 1e10f)) ? 1 : 0;  // false for NaN
        }
        fir151_sq_exact_kernel<<<g, SQ_NT, 0, st>>>(d, make_taps(j.h), td, bounded);
        e = (int)cudaGetLastError();
    } else if (j.decim == 1) e = launch_fir_dk<1, KIND>(j, d, grid, st);
    else if (j.decim == 5 && (KIND == SRC_PLAIN || KIND == SRC_MIX_LATE)) e = launch_fir_dk<5, (KIND == SRC_PLAIN || KIND == SRC_MIX_LATE) ? KIND : SRC_PLAIN>(j, d, grid, st);
    else if (j.decim == 10 && KIND == SRC_PLAIN) e = launch_fir_dk<10, SRC_PLAIN>(j, d, grid, st);
    else return (int)cudaErrorInvalidValue;
    launch_counter() += 1;
    return e;
}

}  // namespace

int launch_fir_multi(const FirMultiJob &j, fmrx_stream_t st) {
    if (j.nf < 2 || j.nf > MAXF || j.n % (j.exact ? 1024 : 1024) != 0 || j.nzi < kHist) return (int)cudaErrorInvalidValue;
    MultiDev d{};
    TapsF t{};
    d.x = j.x; d.ldx = j.ldx; d.ldy = j.ldy; d.nzi = j.nzi; d.n = j.n; d.ny = j.n; d.n_blocks = j.n_blocks; d.nf = j.nf;
    d.zi_first = j.nzi > kHist ? j.nzi - kHist : 0;
    for (int f = 0; f < j.nf; ++f) {
        d.y[f] = j.y[f]; d.zi[f] = j.zi[f];
        for (int k = 0; k < kTaps; ++k) t.h[f][k] = j.h[f][k];
    }
    const Exact2 ex{-0.0f, 1.0f};
    dim3 grid(j.n / 1024, j.n_blocks, j.n_streams);
    if (j.exact) fir151_multi_kernel<true, 128><<<grid, 128, 0, st>>>(d, t, ex);
    else fir151_multi_kernel<false, 64><<<grid, 64, 0, st>>>(d, t, ex);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

int launch_fir(const FirJob &j, fmrx_stream_t st) {
    FirDev d;
    d.x = j.x; d.x2 = j.x2; d.y = j.y; d.zi = j.zi + (j.nzi - kHist); d.zi_out = j.zi;
    d.zi_first = j.live_state_only && j.nzi > kHist ? j.nzi - kHist : 0;
    d.ldx = j.ldx; d.ldy = j.ldy; d.nzi = j.nzi; d.n = j.n; d.ny = j.n / j.decim; d.n_blocks = j.n_blocks;
    constexpr int TO = R * CT;
    dim3 grid((d.ny + TO - 1) / TO, j.n_blocks, j.n_streams);
    switch (j.kind) {
        case SRC_PLAIN: return launch_fir_k<SRC_PLAIN>(j, d, grid, st);
        case SRC_SQUARE: return launch_fir_k<SRC_SQUARE>(j, d, grid, st);
        case SRC_MIX_LATE: return launch_fir_k<SRC_MIX_LATE>(j, d, grid, st);
        case SRC_MIX_HALF: return launch_fir_k<SRC_MIX_HALF>(j, d, grid, st);
        case SRC_PROD_HALF: return launch_fir_k<SRC_PROD_HALF>(j, d, grid, st);
        default: return (int)cudaErrorInvalidValue;
    }
}

template <int D, bool RAW>
static int launch_iq_d(const IqDev &d, const float *h, int exact, int n_streams, fmrx_stream_t st) {
    constexpr int TO = Geom<D, CTQ>::TO;
    const Taps t = make_taps(h);
    dim3 grid((d.ny + TO - 1) / TO, d.n_blocks, n_streams);
    if (exact) fir151_iq_kernel<D, RAW, true><<<grid, CTQ, 0, st>>>(d, t);
    else fir151_iq_kernel<D, RAW, false><<<grid, CTQ, 0, st>>>(d, t);
    cudaError_t e = cudaGetLastError();
    if (e) return (int)e;
    iq_state_kernel<RAW><<<n_streams, 160, 0, st>>>(d);
    launch_counter() += 2;
    return (int)cudaGetLastError();
}

// run length of the streaming front end: a multiple of 4; the largest one <= 240 that divides ny when there is one.
// (Sizing L for whole waves of resident CTAs was tried and buys nothing: the kernel is FP32-pipe bound, so a partly
// filled last wave simply runs its CTAs faster.)
static int stream_run(int ny) {
    for (int L = 240; L >= 64; L -= 4)
        if (ny % L == 0) return L;
    return 240;
}

int launch_fir_iq(const FirIqJob &j, fmrx_stream_t st) {
    IqDev d{};
    d.xi = j.xi; d.xq = j.xq; d.yi = j.yi; d.yq = j.yq; d.zii = j.zii; d.ziq = j.ziq;
    d.ldx = j.ldx; d.ldy = j.ldy; d.n = j.n; d.ny = j.n / j.decim; d.n_blocks = j.n_blocks;
    if (j.decim != 10) return (int)cudaErrorInvalidValue;  // the reference only ever calls it with rf_decim = 10 (src/fm_radio.cpp:42,78)
    return launch_iq_d<10, false>(d, j.h, j.exact, j.n_streams, st);
}

int launch_frontend(const FrontendJob &j, fmrx_stream_t st) {
    IqDev d{};
    d.raw = j.raw; d.demod = j.demod; d.yi = j.yi; d.yq = j.yq; d.zii = j.zii; d.ziq = j.ziq;
    d.ldx = j.ld_raw; d.ldy = j.ld_out; d.n = j.n; d.ny = j.n / 10; d.n_blocks = j.n_blocks;
    RowTaps t;
    for (int jj = 0; jj < 10; ++jj)
        for (int sl = 0; sl < 16; ++sl) {
            const int k = 10 * sl - 1 + jj;
            t.t[jj][sl] = (k >= 0 && k < kTaps) ? j.h[k] : 0.0f;
        }
    const int L = stream_run(d.ny), segs = (d.ny + L - 1) / L;
    const long long total = (long long)segs * j.n_blocks * j.n_streams;
    // the group-walk form needs whole runs, whole quads and 16-byte aligned rows; anything else takes the first form
    static const int forced = [] { const char *v = getenv("FMRX_FRONTEND"); return v ? atoi(v) : 0; }();
    auto al16 = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
    const bool walk4 = forced != 1 && d.n == 10 * d.ny && d.ny % L == 0 && L % 4 == 0 && d.ny >= 32 && al16(d.raw) && d.ldx % 16 == 0 && (2LL * d.n) % 16 == 0 &&
                       al16(d.demod) && al16(d.yi) && al16(d.yq) && d.ldy % 4 == 0 && d.ny % 4 == 0;
    if (walk4) {
        frontend_stream4_kernel<true><<<(unsigned)((total + 63) / 64), 64, 0, st>>>(d, t, L, segs, total, Exact2{-0.0f, 1.0f});
        cudaError_t e4 = cudaGetLastError();
        if (e4) return (int)e4;
        const int n_sb = j.n_blocks * j.n_streams;
        frontend_edge_kernel<true><<<(n_sb + 7) / 8, 128, 0, st>>>(d, make_taps(j.h), n_sb);
        launch_counter() += 1;
    } else {
        frontend_stream_kernel<true><<<(unsigned)((total + 63) / 64), 64, 0, st>>>(d, t, L, segs, total, Exact2{-0.0f, 1.0f});
    }
    cudaError_t e = cudaGetLastError();
    if (e) return (int)e;
    iq_state_kernel<true><<<j.n_streams, 160, 0, st>>>(d);
    launch_counter() += 2;
    return (int)cudaGetLastError();
}

int launch_unpack(const uint8_t *raw, size_t n, float *out, fmrx_stream_t st) {
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    unpack_kernel<<<blocks ? blocks : 1, 256, 0, st>>>(raw, n, out);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

int launch_demod(const float *I, const float *Q, float *out, int n_streams, int n_blocks, int n, fmrx_stream_t st) {
    const long long total = (long long)n_streams * n_blocks * n;
    demod_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(I, Q, out, n, total);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
