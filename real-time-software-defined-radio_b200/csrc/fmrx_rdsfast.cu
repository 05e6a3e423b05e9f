// RDS back end at SYMBOL rate (SURVEY 8a rows a15, a12, a9 as one composite filter; north_star: "compute only the retained
// output phases").
//
// Behind the 114 kHz PLL the reference runs three full-rate filters per block -- mixer + 3 kHz LPF (15360 outputs x 151
// taps, src/filter.cpp:373-401), the 19/80 polyphase resampler (3648 x 151, :301-339) and the RRC matched filter
// (3648 x 151, :126-154) -- and then frame_thread looks at ONE RRC sample in 24 (152 per block, src/fm_radio.cpp:
// 519-526).  All three are linear and time-invariant inside a block, so the sample the decoder reads,
// rrc[24k + off], is a single 933-tap polyphase filter applied to the mixer product p = NCO x RDS band that the PLL kernel
// already writes:  rrc[i] = sum_j W[i mod 19][j] * p[floor(80 i / 19) - j],  W = hr * (19 h2) * (2 h1) laid out per phase.
// 142 of a block's 152 symbols are such interior samples: 0.13 M MACs per block instead of 3.4 M.
//
// What is NOT time-invariant are the reference's block-edge semantics -- the mixer's half-weight history (Q8), the
// resampler's history indexed by tap count (Q6), the RRC's one-late history (Q1) -- and they reach the first ten symbols
// of a block.  Those (and, in the very first block, the 24 samples frame_thread picks its sampling phase from, Q11) are
// computed by a head kernel that restates the three stages exactly as the staged kernels do, on the 1008 + 240 samples
// they need.  The carried state keeps its meaning: the last 150 mixer products, the eight entries of the resampler's
// state its history map can reach, and the last 150 resampler outputs (one late), the latter two evaluated directly from
// p with the composite taps.
//
//   rds_head_kernel    one CTA per (station, block): X window -> rlpf[0..1008) -> rres[0..240) -> rrc at the head symbols
//   rds_symbol_kernel  one CTA per (station, block), one warp per filter phase: the 142 interior symbols; for the last
//                      block of a call also the new carried state, from the same staged samples
// The decoder kernel is unchanged: the symbols are written at their positions 24k + off of the (otherwise untouched) RRC
// buffer.  FMRX_PATH_RDS_STAGES selects the staged kernels instead (every stage materialised, debug taps available).
#include <cuda_runtime.h>

#include <vector>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

constexpr int NIF = FMRX_IF_PER_BLOCK, NRDS = FMRX_RDS_PER_BLOCK, SPS = 24, NSYM = NRDS / SPS;
constexpr int U = 19, D = 80, TPP = kTaps;
constexpr int HEAD_SYMS = 10;                 // symbols 0..9 (rrc index < 240) feel the block edge
constexpr int NRH = HEAD_SYMS * SPS;          // 240 resampler outputs restated by the head kernel
constexpr int NLH = 1008;                     // mixer-LPF outputs they need: floor(80*239/19) = 1006
constexpr int WLEN = 960;                     // composite taps per phase, zero padded (support 933)
constexpr int GLEN = 320;                     // mixer-LPF x resampler composite per phase (support 301), zero padded to 10 x 32
constexpr int ZA_LO = 143, ZA_N = 8;          // entries of the resampler state its history map can reach: (2867 - c)/19, c <= 150

__host__ __device__ constexpr int qof(int o) { return (D * o) / U; }
__host__ __device__ constexpr int phof(int o) { return (D * o) % U; }

struct Taps151 {
    float h[kTaps + 1];
};

struct FastDev {
    const float *p;        // [S][ld] mixer product, n_blocks * NIF per station
    float *rrc;            // [S][ldr] sparse RRC buffer, n_blocks * NRDS per station
    float *zi_lpf, *zi_anti, *zi_rrc;
    const float *h2p;      // device: anti-image taps, phase-major [19][152]: h2p[ph][c] = h2[ph + 19c]
    const float *W, *G;    // device: [19][WLEN], [19][GLEN]
    const int32_t *off;    // sampling phase per station: off[s * off_stride]
    int32_t *off_out;      // where the first block's phase is written (phase_only pass)
    long long ld, ldr;
    int off_stride, n_blocks, first_block_is_zero, nzi_anti;
};

// rres[o] for an interior o of block `pb` (every tap in-block): composite of mixer LPF and resampler, straight from p.
// Warp-cooperative: the 301 taps are split over the lanes (coalesced loads of taps and samples), then reduced.
__device__ __forceinline__ float rres_interior_warp(const float *pb, const float *G, int o, int lane) {
    const float *g = G + phof(o) * GLEN + lane;
    const float *x = pb + qof(o) - lane;
    float acc = 0.0f;
#pragma unroll
    for (int it = 0; it < GLEN / 32; ++it) acc = fmaf(__ldg(g + 32 * it), __ldg(x - 32 * it), acc);  // G is zero beyond its 301 taps
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    return acc;
}

// phase_only: grid (1, S), computes the 24 first RRC samples of block 0 and the sampling phase frame_thread derives from
// them (src/fm_radio.cpp:503-517); otherwise grid (n_blocks, S): head symbols (and those 24 samples again, for the decoder)
template <bool PHASE_ONLY>
__global__ void __launch_bounds__(256) rds_head_kernel(const FastDev a, const __grid_constant__ Taps151 h1, const __grid_constant__ Taps151 hr) {
    __shared__ __align__(16) float xw[kHist + NLH + 10];  // X(j), j = -150 .. NLH-1: history as stored, in-block doubled (+ pad for the last quad)
    __shared__ float rl[NLH];               // rlpf head
    __shared__ float rh[kHist + NRH];       // R(j), j = -150 .. 239: one-late history then rres head
    __shared__ float a8[ZA_N];
    __shared__ float dense[SPS];
    const int b = blockIdx.x, s = blockIdx.y, t = threadIdx.x;
    const float *pb = a.p + (long long)s * a.ld + (long long)b * NIF;
    // ---- X window (Q8: the history holds the product without its x2) and the reachable resampler / RRC history
    for (int i = t; i < kHist + NLH + 10; i += 256) {
        const int j = i - kHist;
        xw[i] = j >= NLH ? 0.0f : j >= 0 ? __fmul_rn(pb[j], 2.0f) : (b > 0 ? pb[j] : a.zi_lpf[(long long)s * kHist + kHist + j]);
    }
    if (b == 0) {
        if (t < ZA_N) a8[t] = a.zi_anti[(long long)s * a.nzi_anti + ZA_LO + t];
    } else {  // rlpf of the previous block at 12492 + 143 + e: interior, straight from p; one warp per entry
        const int e = t >> 5, lane_ = t & 31;
        const float *x = pb - NIF + (NIF + 1 - a.nzi_anti - 1) + ZA_LO + e;
        float acc = 0.0f;
        for (int k = lane_; k < kTaps; k += 32) acc = fmaf(__fmul_rn(x[-k], 2.0f), h1.h[k], acc);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
        if (lane_ == 0) a8[e] = acc;
    }
    // R(-j') = rres_prev[NRDS - 1 - j'] (Q1): entry i = 150 - j' holds rres_prev[NRDS - 151 + i]
    if (b == 0) {
        if (t < kHist) rh[t] = a.zi_rrc[(long long)s * kHist + t];
    } else {
        for (int i = t >> 5; i < kHist; i += 8) {
            const float v = rres_interior_warp(pb - NIF, a.G, NRDS - kHist - 1 + i, t & 31);
            if ((t & 31) == 0) rh[i] = v;
        }
    }
    __syncthreads();
    // ---- mixer LPF on the head (src/filter.cpp:381-396), taps ascending.  Eight consecutive outputs per thread, the
    // loop over the thread's samples newest first (one 128-bit LDS per four, each applied to every output it feeds)
    constexpr int NL = PHASE_ONLY ? (qof(SPS - 1) + 8) / 8 * 8 : NLH;
    if (8 * t < NL) {
        const float *wnd = xw + 8 * t;  // window offset c <-> X(8t - 150 + c); output r uses it with tap k = 150 + r - c
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = 0.0f;
#pragma unroll
        for (int q = (kHist + 7) / 4; q >= 0; --q) {
            const float4 v = *reinterpret_cast<const float4 *>(wnd + 4 * q);
            const float xv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 3; e >= 0; --e) {
                const int c = 4 * q + e;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int k = kHist + r - c;
                    if (k >= 0 && k < kTaps) acc[r] = fmaf(xv[e], h1.h[k], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) rl[8 * t + r] = acc[r];
    }
    __syncthreads();
    // ---- resampler on the head (src/filter.cpp:317-334): in-block taps read rlpf, the others the state at (Z-1-c)/U (Q6)
    constexpr int NR = PHASE_ONLY ? SPS : NRH;
    if (t < NR) {
        const int q0 = qof(t);
        const float4 *hp = reinterpret_cast<const float4 *>(a.h2p + phof(t) * 152);  // phase-major copy: hp[c] = h2[ph + 19c]
        float acc = 0.0f;
        if (q0 >= TPP - 1) {  // every tap in-block (all but the first 36 outputs): no case analysis in the loop
            const float *r = rl + q0;
#pragma unroll 2
            for (int c4 = 0; c4 < 152 / 4; ++c4) {
                const float4 h = __ldg(hp + c4);
                acc = fmaf(r[-4 * c4], h.x, acc);
                acc = fmaf(r[-4 * c4 - 1], h.y, acc);
                acc = fmaf(r[-4 * c4 - 2], h.z, acc);
                if (c4 < 37) acc = fmaf(r[-4 * c4 - 3], h.w, acc);  // tap 151 does not exist
            }
        } else {
            const float *hs = reinterpret_cast<const float *>(hp);
            for (int c = 0; c < TPP; ++c) {
                const float v = c <= q0 ? rl[q0 - c] : a8[(a.nzi_anti - 1 - c) / U - ZA_LO];
                acc = fmaf(v, __ldg(hs + c), acc);
            }
        }
        rh[kHist + t] = __fmul_rn(acc, (float)U);
    }
    __syncthreads();
    // ---- RRC at the positions the decoder reads
    const bool want_dense = PHASE_ONLY || (a.first_block_is_zero && b == 0);
    const int off = PHASE_ONLY ? 0 : a.off[(long long)s * a.off_stride];
    const int n_out = (PHASE_ONLY ? 0 : HEAD_SYMS) + (want_dense ? SPS : 0);
    if (t < n_out) {
        const int i = t < (PHASE_ONLY ? 0 : HEAD_SYMS) ? SPS * t + off : t - (PHASE_ONLY ? 0 : HEAD_SYMS);
        const float *r = rh + kHist + i;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) acc = fmaf(r[-k], hr.h[k], acc);
        if (PHASE_ONLY) dense[i] = acc;
        else a.rrc[(long long)s * a.ldr + (long long)b * NRDS + i] = acc;
    }
    if (PHASE_ONLY) {
        __syncthreads();
        if (t == 0) {  // frame_thread's pick: the first strict maximum of |rrc[0..23]|
            float best = fabsf(dense[0]);
            int o = 0;
            for (int i = 1; i < SPS; ++i)
                if (fabsf(dense[i]) > best) { best = fabsf(dense[i]); o = i; }
            a.off_out[s] = o;
        }
    }
}

// interior symbols: the block's p is staged once in shared memory (asynchronous 16-byte copies); warp w handles filter
// phases w, w+8, w+16, a phase's taps staying in registers (30 per lane) for its 7-8 symbols; each symbol is 30 conflict-free
// LDS + FFMA per lane and a warp reduction
__global__ void __launch_bounds__(256) rds_symbol_kernel(const FastDev a, const __grid_constant__ Taps151 h1) {
    extern __shared__ __align__(16) float ps[];  // NIF floats
    const int b = blockIdx.x, s = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *pb = a.p + (long long)s * a.ld + (long long)b * NIF;
    float *out = a.rrc + (long long)s * a.ldr + (long long)b * NRDS;
    const int off = a.off[(long long)s * a.off_stride];
    if (((uintptr_t)pb & 15) == 0) {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(ps);
        for (int i = threadIdx.x; i < NIF / 4; i += 256) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * i), "l"(pb + 4 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int i = threadIdx.x; i < NIF; i += 256) ps[i] = pb[i];
    }
    __syncthreads();
    for (int phi = warp; phi < U; phi += 8) {
        float w[WLEN / 32];
#pragma unroll
        for (int i = 0; i < WLEN / 32; ++i) w[i] = __ldg(a.W + phi * WLEN + 32 * i + lane);
        // symbols k with (24k + off) mod 19 == phi: 5k = phi - off (mod 19), 5^-1 = 4
        int k = (4 * (((phi - off) % U + U) % U)) % U;
        while (k < HEAD_SYMS) k += U;
        for (; k < NSYM; k += U) {
            const int i = SPS * k + off;
            const float *x = ps + qof(i) - lane;
            float acc = 0.0f;
#pragma unroll
            for (int it = 0; it < WLEN / 32; ++it) acc = fmaf(w[it], x[-32 * it], acc);
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
            if (lane == 0) out[i] = acc;
        }
    }
    if (b != a.n_blocks - 1) return;
    // ---- the last block of the call also leaves the carried state, from the same staged samples
    if (threadIdx.x < kHist) a.zi_lpf[(long long)s * kHist + threadIdx.x] = ps[NIF - kHist + threadIdx.x];  // src/filter.cpp:398-400 at its call site (Q8)
    {   // the eight reachable entries of the resampler state: rlpf[12492 + 143 + e], one warp each
        const float *x = ps + (NIF + 1 - a.nzi_anti - 1) + ZA_LO + warp;
        float acc = 0.0f;
        for (int k = lane; k < kTaps; k += 32) acc = fmaf(__fmul_rn(x[-k], 2.0f), h1.h[k], acc);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
        if (lane == 0) a.zi_anti[(long long)s * a.nzi_anti + ZA_LO + warp] = acc;
    }
    // the last 150 resampler outputs, one late (Q1): rres[3497 + i] = sum_d G[ph][d] p[q - d]; a warp takes a phase, keeps its
    // composite taps in registers and walks the 7-8 outputs of that phase (ph(o) = 4o mod 19, so o = 5 ph mod 19)
    for (int ph = warp; ph < U; ph += 8) {
        float g[GLEN / 32];
#pragma unroll
        for (int it = 0; it < GLEN / 32; ++it) g[it] = __ldg(a.G + ph * GLEN + 32 * it + lane);
        constexpr int O0 = NRDS - kHist - 1;
        int o = O0 + (((5 * ph) % U - O0 % U) % U + U) % U;
        for (; o < O0 + kHist; o += U) {
            const float *x = ps + qof(o) - lane;
            float acc = 0.0f;
#pragma unroll
            for (int it = 0; it < GLEN / 32; ++it) acc = fmaf(g[it], x[-32 * it], acc);
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
            if (lane == 0) a.zi_rrc[(long long)s * kHist + (o - O0)] = acc;
        }
    }
}

struct Composite {
    std::vector<float> W, G;
};

// W[phi][j]: rrc[i] = sum_j W[i%19][j] p[q(i) - j];  G[ph][d]: rres[o] = sum_d G[ph(o)][d] p[q(o) - d]  (double, then fp32)
Composite make_composite(const float *h1, const float *h2, const float *hr) {
    std::vector<double> G((size_t)U * GLEN, 0.0), W((size_t)U * WLEN, 0.0);
    for (int ph = 0; ph < U; ++ph)
        for (int c = 0; c < TPP; ++c)
            for (int k = 0; k < kTaps; ++k) G[(size_t)ph * GLEN + c + k] += (double)U * (double)h2[ph + U * c] * 2.0 * (double)h1[k];
    for (int phi = 0; phi < U; ++phi) {
        const int i = phi + U * 40, M = qof(i);
        for (int aa = 0; aa < kTaps; ++aa) {
            const int o = i - aa, base = M - qof(o);
            const double *g = &G[(size_t)phof(o) * GLEN];
            for (int d = 0; d <= 2 * (kTaps - 1); ++d) W[(size_t)phi * WLEN + base + d] += (double)hr[aa] * g[d];
        }
    }
    Composite c;
    c.W.assign(W.begin(), W.end());
    c.G.assign(G.begin(), G.end());
    return c;
}

Taps151 pack(const float *h) {
    Taps151 t;
    for (int k = 0; k < kTaps; ++k) t.h[k] = h[k];
    t.h[kTaps] = 0.0f;
    return t;
}

}  // namespace

int rds_fast_tables(const float *h1, const float *h2, const float *hr, float **dW, float **dG, float **dH2p) {
    const Composite c = make_composite(h1, h2, hr);
    std::vector<float> h2p((size_t)U * 152, 0.0f);
    for (int ph = 0; ph < U; ++ph)
        for (int cc = 0; cc < TPP; ++cc) h2p[(size_t)ph * 152 + cc] = h2[ph + U * cc];
    if (cudaError_t e0 = cudaMalloc(dH2p, h2p.size() * sizeof(float))) return (int)e0;
    if (cudaError_t e0 = cudaMemcpy(*dH2p, h2p.data(), h2p.size() * sizeof(float), cudaMemcpyHostToDevice)) return (int)e0;
    cudaError_t e = cudaMalloc(dW, c.W.size() * sizeof(float));
    if (e) return (int)e;
    e = cudaMalloc(dG, c.G.size() * sizeof(float));
    if (e) return (int)e;
    e = cudaMemcpy(*dW, c.W.data(), c.W.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e) return (int)e;
    return (int)cudaMemcpy(*dG, c.G.data(), c.G.size() * sizeof(float), cudaMemcpyHostToDevice);
}

int launch_rds_fast(const RdsFastJob &j, fmrx_stream_t st) {
    FastDev d{};
    d.p = j.p; d.rrc = j.rrc; d.zi_lpf = j.zi_lpf; d.zi_anti = j.zi_anti; d.zi_rrc = j.zi_rrc; d.h2p = j.h2p; d.W = j.W; d.G = j.G;
    d.ld = j.ld; d.ldr = j.ldr; d.n_blocks = j.n_blocks; d.first_block_is_zero = j.first_block_is_zero; d.nzi_anti = j.nzi_anti;
    const Taps151 h1 = pack(j.h1), hr = pack(j.hr);
    if (j.first_block_is_zero) {  // the sampling phase does not exist yet: derive it from block 0 first
        d.off_out = j.off_scratch;
        rds_head_kernel<true><<<dim3(1, j.n_streams), 256, 0, st>>>(d, h1, hr);
        d.off = j.off_scratch; d.off_stride = 1;
        launch_counter() += 1;
    } else {
        d.off = j.off_state; d.off_stride = j.off_state_stride;
    }
    rds_head_kernel<false><<<dim3(j.n_blocks, j.n_streams), 256, 0, st>>>(d, h1, hr);
    // per device (per context), so set on every launch: a process may hold handles on several GPUs (fmrx_config.device)
    if (cudaError_t e0 = cudaFuncSetAttribute(rds_symbol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NIF * (int)sizeof(float))) return (int)e0;
    rds_symbol_kernel<<<dim3(j.n_blocks, j.n_streams), 256, NIF * sizeof(float), st>>>(d, h1);
    launch_counter() += 2;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
