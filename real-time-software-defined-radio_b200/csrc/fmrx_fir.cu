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
// B200 mapping: one CTA = 128 compute threads x R=8 consecutive outputs = a 1024-output tile of one (stream, block).
// The input span is staged once in shared memory with a row pitch of d*R+1 words, so that lane t's window starts at
// (d*R+1)*t: odd lane stride = conflict-free LDS, and every tap's offset is a compile-time constant.  The tap loop is
// fully unrolled; taps arrive BY VALUE in the kernel parameter block, so each tap is a constant-bank operand of the
// multiply.  `EXACT` keeps the reference's two roundings per tap (FMUL, FADD); otherwise FFMA.
#include <cuda_runtime.h>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

constexpr int R = 8;          // outputs per thread
constexpr int CT = 128;       // compute threads per CTA
constexpr int TO = R * CT;    // outputs per tile
constexpr int OFF = 168;      // tile origin sits OFF input samples before the first output's newest sample: >= 150 + 10
                              // (history of the output just before the tile, for the discriminator) and 2*OFF % 16 == 0

struct Taps {
    float h[kTaps + 1];
};

template <int D>
struct Geom {
    static constexpr int ROW = D * R;          // input samples between consecutive threads' windows
    static constexpr int PITCH = ROW + 1;      // padded row -> odd lane stride
    static constexpr int SPAN = D * (TO - 1) + OFF + 1;
    static constexpr int WORDS = SPAN + SPAN / ROW + 1;
    __host__ __device__ static constexpr int phys(int i) { return i + i / ROW; }
};

template <bool EXACT>
__device__ __forceinline__ float mac(float acc, float x, float h) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(x, h)) : fmaf(x, h, acc);
}

// ------------------------------------------------------------------------------------------------------------------
// input formation
// ------------------------------------------------------------------------------------------------------------------
struct FirDev {
    const float *x, *x2;
    float *y;
    const float *zi;  // points at the 150 live entries of this launch's state: zi_live[s*nzi + (nzi-150) ...]
    long long ldx, ldy;
    int nzi, n, ny, n_blocks;
};

template <int KIND>
__device__ __forceinline__ float form(float a, float b, bool in_block) {
    if (KIND == SRC_PLAIN) return a;
    if (KIND == SRC_SQUARE) return __fmul_rn(a, a);
    if (KIND == SRC_MIX_LATE) return __fmul_rn(a, b);
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
        q = (long long)b * a.n + p - (KIND == SRC_MIX_HALF ? 0 : 1);  // one-late history, except the mixer (Q8)
    } else {
        return zs[kHist + p];  // carried state already holds the formed value
    }
    float v = xs[q];
    float w = (KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF) ? x2s[q] : 0.0f;
    return form<KIND>(v, w, in_block);
}

// ------------------------------------------------------------------------------------------------------------------
// single-channel kernel
// ------------------------------------------------------------------------------------------------------------------
template <int D, int KIND, bool EXACT>
__global__ void __launch_bounds__(CT) fir151_kernel(const FirDev a, const __grid_constant__ Taps taps) {
    using G = Geom<D>;
    __shared__ float sm[G::WORDS];
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * TO;
    const float *xs = a.x + (long long)s * a.ldx;
    const float *x2s = a.x2 ? a.x2 + (long long)s * a.ldx : nullptr;
    const float *zs = a.zi + (long long)s * a.nzi;
    const int P0 = D * n0 - OFF;
    for (int i = threadIdx.x; i < G::SPAN; i += CT) sm[G::phys(i)] = source<KIND>(a, xs, x2s, zs, b, P0 + i);
    __syncthreads();

    const float *w = sm + G::PITCH * threadIdx.x;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.0f;
#pragma unroll
    for (int k = 0; k < kTaps; ++k) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int c = D * r - k + OFF;
            acc[r] = mac<EXACT>(acc[r], w[c + c / G::ROW], taps.h[k]);
        }
    }
    const int o = n0 + R * threadIdx.x;
    float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny + o;
    if (o + R <= a.ny && ((reinterpret_cast<uintptr_t>(ys) & 15) == 0)) {
        reinterpret_cast<float4 *>(ys)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4 *>(ys)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (o + r < a.ny) ys[r] = acc[r];
    }
}

// state after the launch: zi[i] = formed(x[N - nzi - 1 + i]) of the LAST block (mixer: N - nzi + i, no x2)
template <int KIND>
__global__ void fir_state_kernel(const float *x, const float *x2, float *zi, long long ldx, int nzi, int n, int n_blocks) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nzi) return;
    const int p = n - nzi + i - (KIND == SRC_MIX_HALF ? 0 : 1);
    if (p < 0) return;  // block shorter than the state: entry keeps its old value (never live for a 151-tap filter)
    const long long q = (long long)s * ldx + (long long)(n_blocks - 1) * n + p;
    const float v = x[q];
    const float w = (KIND == SRC_MIX_LATE || KIND == SRC_MIX_HALF) ? x2[q] : 0.0f;
    zi[(long long)s * nzi + i] = form<KIND>(v, w, false);
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
__global__ void __launch_bounds__(CT + 32) fir151_iq_kernel(const IqDev a, const __grid_constant__ Taps taps) {
    using G = Geom<D>;
    extern __shared__ float2 smq[];
    __shared__ float2 edge[CT + 1];  // edge[t+1] = last output of thread t; edge[0] = output just before the tile
    const int s = blockIdx.z, b = blockIdx.y, n0 = blockIdx.x * TO;
    const int P0 = D * n0 - OFF;
    if (RAW && (((uintptr_t)a.raw | (uintptr_t)a.ldx | (uintptr_t)(2 * a.n)) & 3) == 0) {
        // Fast ingest: two complex samples (4 bytes) per lane per load, every load of the tile issued before the first
        // one is consumed (memory-level parallelism PAIRS/160 = 33 per thread), then unpack + STS.  Pairs that touch
        // the history (p < 0) or the block end go through the generic per-sample path.
        constexpr int PAIRS = (G::SPAN + 1) / 2, NT = CT + 32, ITER = (PAIRS + NT - 1) / NT;
        const uint8_t *row = a.raw + (long long)s * a.ldx + 2LL * b * a.n;
        unsigned v[ITER];
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int p = P0 + 2 * (threadIdx.x + it * NT);
            v[it] = (p >= 0 && p + 1 < a.n && threadIdx.x + it * NT < PAIRS) ? __ldg(reinterpret_cast<const unsigned *>(row + 2 * p)) : 0u;
        }
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int j = threadIdx.x + it * NT, i = 2 * j, p = P0 + i;
            if (j >= PAIRS) continue;
            if (p >= 0 && p + 1 < a.n) {
                smq[G::phys(i)] = make_float2(u8_lane_to_f32<0>(v[it]), u8_lane_to_f32<1>(v[it]));
                if (i + 1 < G::SPAN) smq[G::phys(i + 1)] = make_float2(u8_lane_to_f32<2>(v[it]), u8_lane_to_f32<3>(v[it]));
            } else {
                smq[G::phys(i)] = source_iq<RAW>(a, s, b, p);
                if (i + 1 < G::SPAN) smq[G::phys(i + 1)] = source_iq<RAW>(a, s, b, p + 1);
            }
        }
    } else {
        for (int i = threadIdx.x; i < G::SPAN; i += CT + 32) smq[G::phys(i)] = source_iq<RAW>(a, s, b, P0 + i);
    }
    __syncthreads();

    float ai[R], aq[R];
    if (threadIdx.x < CT) {
        const float2 *w = smq + G::PITCH * threadIdx.x;
#pragma unroll
        for (int r = 0; r < R; ++r) ai[r] = aq[r] = 0.0f;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int c = D * r - k + OFF;
                const float2 v = w[c + c / G::ROW];
                ai[r] = mac<EXACT>(ai[r], v.x, taps.h[k]);
                aq[r] = mac<EXACT>(aq[r], v.y, taps.h[k]);
            }
        }
        edge[threadIdx.x + 1] = make_float2(ai[R - 1], aq[R - 1]);
    } else if (RAW && threadIdx.x == CT) {
        // the discriminator of the tile's first output needs the filtered sample just before the tile (Q3: zero at a
        // block start).  One lane of the spare warp recomputes it with the same rounding sequence.
        float ei = 0.0f, eq = 0.0f;
        if (n0 > 0) {
            for (int k = 0; k < kTaps; ++k) {
                const int c = -D - k + OFF;  // output n0-1
                const float2 v = smq[G::phys(c)];
                ei = mac<EXACT>(ei, v.x, taps.h[k]);
                eq = mac<EXACT>(eq, v.y, taps.h[k]);
            }
        }
        edge[0] = make_float2(ei, eq);
    }
    if (!RAW && threadIdx.x >= CT) return;
    if (RAW) __syncthreads();
    if (threadIdx.x >= CT) return;

    const int o = n0 + R * threadIdx.x;
    const long long base = (long long)s * a.ldy + (long long)b * a.ny + o;
    if (a.yi) {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (o + r < a.ny) { a.yi[base + r] = ai[r]; a.yq[base + r] = aq[r]; }
    }
    if (RAW) {
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

template <int D, int KIND>
int launch_fir_dk(const FirJob &j, const FirDev &d, dim3 grid, fmrx_stream_t st) {
    const Taps t = make_taps(j.h);
    if (j.exact) fir151_kernel<D, KIND, true><<<grid, CT, 0, st>>>(d, t);
    else fir151_kernel<D, KIND, false><<<grid, CT, 0, st>>>(d, t);
    return (int)cudaGetLastError();
}

template <int KIND>
int launch_fir_k(const FirJob &j, const FirDev &d, dim3 grid, fmrx_stream_t st) {
    int e;
    // only the (kind, decimation) pairs the receive chain and the function-level API use are instantiated
    if (j.decim == 1) e = launch_fir_dk<1, KIND>(j, d, grid, st);
    else if (j.decim == 5 && (KIND == SRC_PLAIN || KIND == SRC_MIX_LATE)) e = launch_fir_dk<5, (KIND == SRC_PLAIN || KIND == SRC_MIX_LATE) ? KIND : SRC_PLAIN>(j, d, grid, st);
    else if (j.decim == 10 && KIND == SRC_PLAIN) e = launch_fir_dk<10, SRC_PLAIN>(j, d, grid, st);
    else return (int)cudaErrorInvalidValue;
    if (e) return e;
    dim3 sg((j.nzi + 127) / 128, j.n_streams);
    fir_state_kernel<KIND><<<sg, 128, 0, st>>>(j.x, j.x2, j.zi, j.ldx, j.nzi, j.n, j.n_blocks);
    launch_counter() += 2;
    return (int)cudaGetLastError();
}

}  // namespace

int launch_fir(const FirJob &j, fmrx_stream_t st) {
    FirDev d;
    d.x = j.x; d.x2 = j.x2; d.y = j.y; d.zi = j.zi + (j.nzi - kHist);
    d.ldx = j.ldx; d.ldy = j.ldy; d.nzi = j.nzi; d.n = j.n; d.ny = j.n / j.decim; d.n_blocks = j.n_blocks;
    dim3 grid((d.ny + TO - 1) / TO, j.n_blocks, j.n_streams);
    switch (j.kind) {
        case SRC_PLAIN: return launch_fir_k<SRC_PLAIN>(j, d, grid, st);
        case SRC_SQUARE: return launch_fir_k<SRC_SQUARE>(j, d, grid, st);
        case SRC_MIX_LATE: return launch_fir_k<SRC_MIX_LATE>(j, d, grid, st);
        case SRC_MIX_HALF: return launch_fir_k<SRC_MIX_HALF>(j, d, grid, st);
        default: return (int)cudaErrorInvalidValue;
    }
}

template <int D, bool RAW>
static int launch_iq_d(const IqDev &d, const float *h, int exact, int n_streams, fmrx_stream_t st) {
    using G = Geom<D>;
    const Taps t = make_taps(h);
    const size_t smem = sizeof(float2) * G::WORDS;
    dim3 grid((d.ny + TO - 1) / TO, d.n_blocks, n_streams);
    cudaError_t e;
    if (exact) {
        e = cudaFuncSetAttribute(fir151_iq_kernel<D, RAW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        fir151_iq_kernel<D, RAW, true><<<grid, CT + 32, smem, st>>>(d, t);
    } else {
        e = cudaFuncSetAttribute(fir151_iq_kernel<D, RAW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e) return (int)e;
        fir151_iq_kernel<D, RAW, false><<<grid, CT + 32, smem, st>>>(d, t);
    }
    e = cudaGetLastError();
    if (e) return (int)e;
    iq_state_kernel<RAW><<<n_streams, 160, 0, st>>>(d);
    launch_counter() += 2;
    return (int)cudaGetLastError();
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
    return launch_iq_d<10, true>(d, j.h, 1, j.n_streams, st);
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
