// RDS back end at SYMBOL rate (SURVEY 8a rows a15, a12, a9 as one composite filter; north_star: "compute only the retained
// output phases").
//
// Behind the 114 kHz PLL the reference runs three full-rate filters per block -- mixer + 3 kHz LPF (15360 outputs x 151
// taps, src/filter.cpp:373-401), the 19/80 polyphase resampler (3648 x 151, :301-339) and the RRC matched filter
// (3648 x 151, :126-154) -- and then frame_thread looks at ONE RRC sample in 24 (152 per block, src/fm_radio.cpp:
// 519-526).  All three are linear, so the sample the decoder reads is one dot product with the mixer product p = NCO x RDS
// band that the PLL kernel already writes:
//
//   interior samples   rrc[i] = sum_j W[i mod 19][j] * p[floor(80 i / 19) - j],  W = hr * (19 h2) * (2 h1) per phase, 933 taps.
//
//   block edge         What is NOT time-invariant are the reference's block-edge semantics -- the mixer's half-weight history
//                      (Q8), the resampler's history indexed by tap count (Q6), the RRC's one-late history (Q1).  They reach
//                      rrc[0..239] (the first ten symbols of a block), and they are still LINEAR: every history value is
//                      itself a fixed combination of the previous block's products (the mixer history IS the last 150 of
//                      them; the eight resampler-state entries the history map can reach are mixer-filter outputs at
//                      12635..12642; the RRC history is the resampler's outputs 3497..3646).  So
//                          rrc[i] = sum_j W[i mod 19][j] * p_ze[q(i) - j]  +  sum_m E[i][m] * p_prev[window(m)],   i < 240,
//                      with p_ze = this block's products preceded by zeros (the causal cascade of a zero-extended input) and
//                      E a 240 x 1094 table over two windows of the previous block, p_prev[12485..12642] and
//                      p_prev[14424..15359], built once on the host in double (make_tables).  Checked against the staged
//                      kernels and the oracle by tests/test_gpu_chain.py::test_rds_symbol_rate_path_equals_staged.
//
// One kernel, one CTA per (station, block): the block's products staged in shared memory behind 960 zeros, the two windows of
// the previous block (read from the previous block of the same call, or from the carried state at the start of a call) next
// to them; a warp takes a filter phase, keeps its 960 taps in registers and walks the 8 symbols of that phase.  The first
// version restated the three stages on the head of the block (1008 + 240 + 10 filter outputs per block, barrier-separated,
// 8.6 M shared-memory bank conflicts per launch: 0.20 ms) and recomputed the three filter states at every block end
// (150 x 301 + 8 x 151 MACs); this one does 152 x 933 + 10 x 1094 MACs per block and keeps 1094 products as its state.
// The decoder kernel is unchanged: the symbols are written at their positions 24k + off of the (otherwise untouched) RRC
// buffer.  FMRX_PATH_RDS_STAGES selects the staged kernels instead (every stage materialised, debug taps available).
#include <cuda_runtime.h>

#include <vector>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

constexpr int NIF = FMRX_IF_PER_BLOCK, NRDS = FMRX_RDS_PER_BLOCK, SPS = 24, NSYM = NRDS / SPS;
constexpr int U = 19, D = 80, TPP = kTaps;
constexpr int HEAD = 240;                     // rrc positions 0..239 (symbols 0..9) feel the block edge
constexpr int WLEN = 960;                     // composite taps per phase, zero padded (support 933): 30 per lane
constexpr int GLEN = 301;                     // mixer-LPF x resampler composite per phase
constexpr int A0 = 12485, AN = 158;           // window A of the previous block: feeds the 8 reachable resampler-state entries
constexpr int B0 = 14424, BN = NIF - B0;      // window B: feeds the RRC history (resampler outputs 3497..3646) and the mixer history
constexpr int EN = AN + BN;                   // 1094 edge products
constexpr int ELEN = 1120;                    // padded to 35 per lane
constexpr int ZPAD = WLEN;                    // zeros staged ahead of the block
static_assert(BN == 936 && EN == 1094 && ELEN % 32 == 0 && ELEN >= EN, "edge window geometry");

__host__ __device__ constexpr int qof(int o) { return (D * o) / U; }
__host__ __device__ constexpr int phof(int o) { return (D * o) % U; }

struct FastDev {
    const float *p;        // [S][ld] mixer product, n_blocks * NIF per station
    float *rrc;            // [S][ldr] sparse RRC buffer, n_blocks * NRDS per station
    float *edge;           // [S][edge_stride] carried state: the EN edge products of the last block processed
    const float *W, *E;    // device: [19][WLEN], [HEAD][ELEN]
    const int32_t *off;    // sampling phase per station: off[s * off_stride]
    int32_t *off_out;      // where the first block's phase is written (phase-only pass)
    long long ld, ldr;
    int off_stride, n_blocks, first_block_is_zero, edge_stride;
};

// the very first block of a station: frame_thread picks its sampling phase from |rrc[0..23]| (src/fm_radio.cpp:503-517), so
// those 24 samples come first: one CTA per station, a warp per position
__global__ void __launch_bounds__(256) rds_phase_kernel(const FastDev a) {
    __shared__ float dense[SPS];
    const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *pb = a.p + (long long)s * a.ld;
    const float *ed = a.edge + (long long)s * a.edge_stride;
    for (int i = warp; i < SPS; i += 8) {
        const float *w = a.W + (i % U) * WLEN;
        const int q = qof(i);
        float acc = 0.0f;
        for (int j = lane; j <= q; j += 32) acc = fmaf(__ldg(w + j), pb[q - j], acc);
        const float *e = a.E + (long long)i * ELEN;
        for (int m = lane; m < EN; m += 32) acc = fmaf(__ldg(e + m), ed[m], acc);
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
        if (lane == 0) dense[i] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // the first strict maximum of |rrc[0..23]|
        float best = fabsf(dense[0]);
        int o = 0;
        for (int i = 1; i < SPS; ++i)
            if (fabsf(dense[i]) > best) { best = fabsf(dense[i]); o = i; }
        a.off_out[s] = o;
    }
}

// Ten warps per CTA: the 19 filter phases are dealt out two to a warp (one warp takes one).  With eight warps three of them took three
// phases and the CTA waited for those: 0.199 ms per 4096-station step.
constexpr int SYM_NT = 320;
__global__ void __launch_bounds__(SYM_NT) rds_symbol_kernel(const FastDev a) {
    extern __shared__ __align__(16) float ps[];  // ZPAD zeros | NIF products of this block | ELEN edge products of the previous one
    float *pz = ps + ZPAD, *pw = ps + ZPAD + NIF;
    const int b = blockIdx.x, s = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *pb = a.p + (long long)s * a.ld + (long long)b * NIF;
    float *out = a.rrc + (long long)s * a.ldr + (long long)b * NRDS;
    float *ed = a.edge + (long long)s * a.edge_stride;
    const int off = a.off[(long long)s * a.off_stride];
    const bool aligned = ((uintptr_t)pb & 15) == 0;
    if (aligned) {
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(pz);
        for (int i = threadIdx.x; i < NIF / 4; i += SYM_NT) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + 16u * i), "l"(pb + 4 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        for (int i = threadIdx.x; i < NIF; i += SYM_NT) pz[i] = pb[i];
    }
    for (int i = threadIdx.x; i < ZPAD; i += SYM_NT) ps[i] = 0.0f;
    for (int m = threadIdx.x; m < ELEN; m += SYM_NT) {
        float v = 0.0f;
        if (m < EN) {
            const int src = m < AN ? A0 + m : B0 + (m - AN);
            v = b > 0 ? pb[src - NIF] : ed[m];
        }
        pw[m] = v;
    }
    if (aligned) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const bool dense = a.first_block_is_zero && b == 0;  // the decoder re-derives the sampling phase from rrc[0..23] of block 0
    for (int phi = warp; phi < U; phi += SYM_NT / 32) {
        float w[WLEN / 32];
#pragma unroll
        for (int i = 0; i < WLEN / 32; ++i) w[i] = __ldg(a.W + phi * WLEN + 32 * i + lane);
        auto sample = [&](int i) {
            const float *x = pz + qof(i) - lane;
            float acc = 0.0f;
#pragma unroll
            for (int it = 0; it < WLEN / 32; ++it) acc = fmaf(w[it], x[-32 * it], acc);
            if (i < HEAD) {  // warp-uniform
                const float *e = a.E + (long long)i * ELEN + lane;
#pragma unroll 5
                for (int t = 0; t < ELEN / 32; ++t) acc = fmaf(__ldg(e + 32 * t), pw[32 * t + lane], acc);
            }
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
            if (lane == 0) out[i] = acc;
        };
        // symbols k with (24k + off) mod 19 == phi: 5k = phi - off (mod 19), 5^-1 = 4
        for (int k = (4 * (((phi - off) % U + U) % U)) % U; k < NSYM; k += U) sample(SPS * k + off);
        if (dense)
            for (int i = phi; i < SPS; i += U)
                if (i != off) sample(i);
    }
    if (b != 0) return;
    // the carried state for the next call: the edge products of the LAST block of this one.  Written by the CTA that read the old
    // state (block 0's, after the barrier above), so no other CTA of this launch races with it.
    const float *pl = a.p + (long long)s * a.ld + (long long)(a.n_blocks - 1) * NIF;
    for (int m = threadIdx.x; m < EN; m += SYM_NT) ed[m] = pl[m < AN ? A0 + m : B0 + (m - AN)];
}

struct Tables {
    std::vector<float> W, E;
};

// W[phi][j]: rrc[i] = sum_j W[i%19][j] p[q(i) - j] away from the block edge.
// E[i][m], i < 240: what the previous block's edge products add to rrc[i] through the three histories (see the header).
Tables make_tables(const float *h1, const float *h2, const float *hr) {
    std::vector<double> G((size_t)U * GLEN, 0.0), W((size_t)U * WLEN, 0.0);
    for (int ph = 0; ph < U; ++ph)
        for (int c = 0; c < TPP; ++c)
            for (int k = 0; k < kTaps; ++k) G[(size_t)ph * GLEN + c + k] += (double)U * (double)h2[ph + U * c] * 2.0 * (double)h1[k];
    for (int phi = 0; phi < U; ++phi) {
        const int i = phi + U * 40, M = qof(i);
        for (int aa = 0; aa < kTaps; ++aa) {
            const int o = i - aa, base = M - qof(o);
            const double *g = &G[(size_t)phof(o) * GLEN];
            for (int d = 0; d < GLEN; ++d) W[(size_t)phi * WLEN + base + d] += (double)hr[aa] * g[d];
        }
    }
    // R[j], j = -150 .. 239: coefficients over the edge products of the resampler output the RRC reads at position j
    //   j < 0: the RRC's one-late history (Q1), R(j) = rres_prev[3647 + j], an interior output of the previous block: G over window B
    //   j >= 0: rres[j] of this block, its history part only -- tap c reads the mixer filter at q(j) - c, whose taps k > q(j) - c reach
    //           into the mixer history (the previous block's last products at weight 1 instead of 2, Q8), or, for c > q(j), the
    //           resampler state entry (2867 - c) / 19 (Q6) = the previous block's mixer filter output at 12492 + that index
    std::vector<double> R((size_t)(kHist + HEAD) * EN, 0.0);
    auto row = [&](int j) { return &R[(size_t)(j + kHist) * EN]; };
    for (int j = -kHist; j < 0; ++j) {
        const int o = NRDS - 1 + j, q = qof(o);
        const double *g = &G[(size_t)phof(o) * GLEN];
        for (int d = 0; d < GLEN; ++d) row(j)[AN + (q - d - B0)] += g[d];
    }
    for (int j = 0; j < HEAD; ++j) {
        const int q = qof(j), f = phof(j);
        for (int c = 0; c < TPP; ++c) {
            const double w = (double)U * (double)h2[f + U * c];
            if (c <= q) {
                const int n = q - c;
                for (int k = n + 1; k < kTaps; ++k) row(j)[AN + (NIF + n - k - B0)] += w * (double)h1[k];
            } else {
                const int n = (NIF + 1 - (kTaps * U - 1) - 1) + (kTaps * U - 2 - c) / U;  // 12492 + (2867 - c) / 19
                for (int k = 0; k < kTaps; ++k) row(j)[n - k - A0] += w * 2.0 * (double)h1[k];
            }
        }
    }
    std::vector<double> E((size_t)HEAD * ELEN, 0.0);
    for (int i = 0; i < HEAD; ++i)
        for (int aa = 0; aa < kTaps; ++aa) {
            const double *r = row(i - aa);
            double *e = &E[(size_t)i * ELEN];
            for (int m = 0; m < EN; ++m) e[m] += (double)hr[aa] * r[m];
        }
    Tables t;
    t.W.assign(W.begin(), W.end());
    t.E.assign(E.begin(), E.end());
    return t;
}

}  // namespace

int rds_fast_tables(const float *h1, const float *h2, const float *hr, float **dW, float **dE) {
    const Tables t = make_tables(h1, h2, hr);
    cudaError_t e = cudaMalloc(dW, t.W.size() * sizeof(float));
    if (e) return (int)e;
    e = cudaMalloc(dE, t.E.size() * sizeof(float));
    if (e) return (int)e;
    e = cudaMemcpy(*dW, t.W.data(), t.W.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e) return (int)e;
    return (int)cudaMemcpy(*dE, t.E.data(), t.E.size() * sizeof(float), cudaMemcpyHostToDevice);
}

int launch_rds_fast(const RdsFastJob &j, fmrx_stream_t st) {
    if (j.edge_stride < EN) return (int)cudaErrorInvalidValue;
    FastDev d{};
    d.p = j.p; d.rrc = j.rrc; d.edge = j.edge; d.W = j.W; d.E = j.E;
    d.ld = j.ld; d.ldr = j.ldr; d.n_blocks = j.n_blocks; d.first_block_is_zero = j.first_block_is_zero; d.edge_stride = j.edge_stride;
    if (j.first_block_is_zero) {  // the sampling phase does not exist yet: derive it from block 0 first
        d.off_out = j.off_scratch;
        rds_phase_kernel<<<j.n_streams, 256, 0, st>>>(d);
        d.off = j.off_scratch; d.off_stride = 1;
        launch_counter() += 1;
    } else {
        d.off = j.off_state; d.off_stride = j.off_state_stride;
    }
    const int smem = (ZPAD + NIF + ELEN) * (int)sizeof(float);
    // per device (per context), so set on every launch: a process may hold handles on several GPUs (fmrx_config.device)
    if (cudaError_t e0 = cudaFuncSetAttribute(rds_symbol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) return (int)e0;
    rds_symbol_kernel<<<dim3(j.n_blocks, j.n_streams), SYM_NT, smem, st>>>(d);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
