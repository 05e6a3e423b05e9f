// Rational polyphase resampler and the RDS clock/data recovery + frame synchroniser
// (rows a17-a20).
//
// Reference (file:line under /root/reference):
//   src/filter.cpp:222-339  convolveWithDecimMode1 / ...Pointer / ...RDS: for output o only the taps k = k0 + c*U with
//                           k0 = (D*o) mod U are visited; in-range taps read x[(D*o-k)/U]; the others read
//                           zi[(Z-1-c)/U] where c counts every visited tap (Q6); the RDS variant scales by U (:333);
//                           zi[i] = x[N-Z-1+i] afterwards.  Only the retained phases are computed.
//   src/fm_radio.cpp:444-729 frame_thread: one-shot sampling phase, Manchester alignment screening, biphase decode,
//                           differential decode, sliding 26-bit syndrome against A/B/C/D with the false-positive
//                           counter / resync, 27-bit carry.  The decoder itself is integer logic on COMPARISONS of RRC samples:
//                           given the reference's RRC samples its bits and events are the reference's.  Whether the samples
//                           are the reference's is decided upstream: bit for bit under STRICT numerics with the staged back
//                           end, to 2e-7 with the symbol-rate back end, to 1e-5 .. 1e-7 under REFERENCE numerics (FFMA in
//                           pllCombine's filter) -- there a decision can differ only where two compared samples are closer than
//                           that; measured under AWGN down to 2 dB CNR: 0 of 3036 bits differ in any setting
//                           (tests/test_gpu_chain.py::test_rds_on_noisy_input_agreement).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

struct ResDev {
    const float *x;
    float *y;
    float *zi;
    const float *h;
    long long ldx, ldy;
    int n, n_ref, ny, n_blocks, ntaps, nzi, decim, up, gain_up;
};

template <bool EXACT>
__global__ void resample_kernel(const ResDev a, int o_count) {  // outputs o < o_count of every block (o_count = a.ny: all of them)
    const int s = blockIdx.z, b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= o_count) return;
    const float *xs = a.x + (long long)s * a.ldx;
    const float *zs = a.zi + (long long)s * a.nzi;
    const long long base = (long long)a.decim * o;
    const int k0 = (int)(base % a.up);
    const int q0 = (int)((base - k0) / a.up);
    float acc = 0.0f;
    int c = 0;
    for (int k = k0; k < a.ntaps; k += a.up, ++c) {
        float v;
        if (c <= q0) {
            v = xs[(long long)b * a.n + (q0 - c)];
        } else {
            const int j = (a.nzi - 1 - c) / a.up;
            v = b > 0 ? xs[(long long)(b - 1) * a.n + (a.n_ref - a.nzi - 1 + j)] : zs[j];
        }
        acc = EXACT ? __fadd_rn(acc, __fmul_rn(v, a.h[k])) : fmaf(v, a.h[k], acc);
    }
    if (a.gain_up) acc = __fmul_rn(acc, (float)a.up);
    a.y[(long long)s * a.ldy + (long long)b * a.ny + o] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// The audio resamplers of modes 1 and 2 (x24 /125, x24 /5 [Q14], x147 /800): same arithmetic, grouped by PHASE.
//
// The general kernel above gives a warp 32 consecutive outputs = 32 different phases: every tap load is a 32-line
// gather (taps k0 + U*c with a different k0 per lane) and the kernel runs at ~10 % of the FIR kernels' rate (mode 1:
// 1.05 + 1.34 ms per step against 0.15 ms for the same MAC count in mode 0).  Outputs o = r + U*m (same residue r) share
// their phase (D*o mod U = D*r mod U), so here a warp takes one residue and 32 values of m: the 151 taps of that phase
// are ONE address per load for the whole warp (phase-major table, broadcast), and the samples x[q_r + D*m - c] are read
// from a copy of the block staged in shared memory with lane stride D -- conflict-free when D is odd; an even D gets one
// pad word per D samples (pitch D+1), and because q0 mod D = q_r is the same for all lanes the pad is crossed at a
// warp-uniform tap count, so the loop just splits there.  One CTA per (station, block).  Tasks whose lanes reach into
// the history (q0 < 150: the first outputs of a block) run a second form of the loop that selects, per tap, between the
// staged block and the history value zi[(Z-1-c)/U] -- which depends on the tap count only (Q6), so it is staged once per
// CTA as 151 floats.  Summation order per output is the reference's (c ascending); EXACT keeps FMUL + FADD.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TPP_ = kTaps;           // taps per phase
constexpr int HPITCH = kTaps + 1;     // phase-major table pitch

// the exact tap on a pair of lanes, as in csrc/fmrx_fir.cu: fma(x, h, -0) is the rounded product and fma(p, 1, acc) the
// rounded sum; the -0 and the 1 are kernel parameters so that ptxas cannot fold the pair into one FFMA2 (one rounding)
struct PairConst {
    float negzero, one;
};
template <bool EXACT>
__device__ __forceinline__ float2 pair_mac(float2 acc, float2 x, float h, const PairConst &k) {
    if (EXACT) {
        const float2 p = __ffma2_rn(x, make_float2(h, h), make_float2(k.negzero, k.negzero));
        return __ffma2_rn(p, make_float2(k.one, k.one), acc);
    }
    return __ffma2_rn(x, make_float2(h, h), acc);
}

template <bool EXACT, bool SKEW>
__global__ void __launch_bounds__(256) resample_phase_kernel(const ResDev a, const float *__restrict__ hp, const PairConst pk, int o_skip) {
    // o_skip: outputs below it are left to the general kernel (the launcher sets it to the number of outputs that reach
    // into the history when there are few of them, so that every task here runs the all-in-block loop)
    constexpr int skew = SKEW ? 1 : 0;
    extern __shared__ float xs_sh[];                       // block b of the stream, index p + skew * (p / D)
    __shared__ __align__(16) float hist[TPP_ + 1];         // history value for tap count c (Q6)
    __shared__ __align__(16) float wtaps[8][HPITCH];       // the current task's taps, per warp
    const int b = blockIdx.x, s = blockIdx.y, U = a.up, D = a.decim;
    const float *xb = a.x + (long long)s * a.ldx + (long long)b * a.n;
    for (int p4 = threadIdx.x; p4 < a.n / 4; p4 += 256) {  // n % 4 == 0 and 16-byte alignment checked by the launcher
        const float4 v = __ldg(reinterpret_cast<const float4 *>(xb) + p4);
        const int p = 4 * p4;
        if (skew) {
            xs_sh[p + p / D] = v.x; xs_sh[p + 1 + (p + 1) / D] = v.y; xs_sh[p + 2 + (p + 2) / D] = v.z; xs_sh[p + 3 + (p + 3) / D] = v.w;
        } else {
            *reinterpret_cast<float4 *>(xs_sh + p) = v;
        }
    }
    if (threadIdx.x < TPP_) {
        const int j = (a.nzi - 1 - (int)threadIdx.x) / U;
        hist[threadIdx.x] = b > 0 ? xb[-(long long)a.n + (a.n_ref - a.nzi - 1 + j)] : a.zi[(long long)s * a.nzi + j];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunks = ((a.ny + U - 1) / U + 31) / 32;     // 32 values of m per task
    float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny;
    float *wt = wtaps[warp];
    if (!SKEW) {
        // odd D (modes 1): a lane owns TWO outputs of the residue, m and m + 32, and runs them as one packed pair -- half
        // the FP32 instructions and half the tap loads per MAC; the lane stride inside each half stays D
        const int chunks2 = ((a.ny + U - 1) / U + 63) / 64;
        const int qmin = TPP_ - 1;
        for (int task = warp; task < U * chunks2; task += 8) {
            const int r = task / chunks2, ma = 64 * (task % chunks2) + lane, mb = ma + 32;
            const int oa = r + U * ma, ob = r + U * mb;
            const int ph = (D * r) % U, qr = (D * r) / U;
            const bool va = oa < a.ny && oa >= o_skip, vb = ob < a.ny && ob >= o_skip;
            const int q0a = va ? qr + D * ma : 0x3fffffff, q0b = vb ? qr + D * mb : 0x3fffffff;
            __syncwarp();
            for (int i2 = lane; i2 < HPITCH; i2 += 32) wt[i2] = __ldg(hp + (long long)ph * HPITCH + i2);
            __syncwarp();
            float2 acc = make_float2(0.0f, 0.0f);
            if (__all_sync(0xffffffffu, q0a >= qmin && q0b >= qmin)) {
                const float *xa = xs_sh + (va ? q0a : TPP_), *xb2 = xs_sh + (vb ? q0b : TPP_);
#pragma unroll 2
                for (int c = 0; c + 4 <= TPP_; c += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(wt + c);
                    acc = pair_mac<EXACT>(acc, make_float2(xa[-c], xb2[-c]), t.x, pk);
                    acc = pair_mac<EXACT>(acc, make_float2(xa[-c - 1], xb2[-c - 1]), t.y, pk);
                    acc = pair_mac<EXACT>(acc, make_float2(xa[-c - 2], xb2[-c - 2]), t.z, pk);
                    acc = pair_mac<EXACT>(acc, make_float2(xa[-c - 3], xb2[-c - 3]), t.w, pk);
                }
#pragma unroll
                for (int c = TPP_ / 4 * 4; c < TPP_; ++c) acc = pair_mac<EXACT>(acc, make_float2(xa[-c], xb2[-c]), wt[c], pk);
            } else {                                        // some lane reaches into the history: branch-free select per tap
                auto tap2 = [&](int c, float t, float hv) {
                    const int da = q0a - c, db = q0b - c;   // >= 0: inside the block
                    const float xva = xs_sh[da >= 0 ? da : 0], xvb = xs_sh[db >= 0 ? db : 0];
                    acc = pair_mac<EXACT>(acc, make_float2(da >= 0 ? xva : hv, db >= 0 ? xvb : hv), t, pk);
                };
#pragma unroll 2
                for (int c = 0; c + 4 <= TPP_; c += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(wt + c), hv = *reinterpret_cast<const float4 *>(hist + c);
                    tap2(c, t.x, hv.x); tap2(c + 1, t.y, hv.y); tap2(c + 2, t.z, hv.z); tap2(c + 3, t.w, hv.w);
                }
#pragma unroll
                for (int c = TPP_ / 4 * 4; c < TPP_; ++c) tap2(c, wt[c], hist[c]);
            }
            if (a.gain_up) { acc.x = __fmul_rn(acc.x, (float)U); acc.y = __fmul_rn(acc.y, (float)U); }
            if (va) ys[oa] = acc.x;
            if (vb) ys[ob] = acc.y;
        }
        return;
    }
    for (int task = warp; task < U * chunks; task += 8) {
        const int r = task / chunks, m = 32 * (task % chunks) + lane;
        const int o = r + U * m;
        const int ph = (D * r) % U, qr = (D * r) / U;       // q0 = qr + D*m, qr < D
        const bool valid = o < a.ny && o >= o_skip;
        const int q0 = valid ? qr + D * m : 0x3fffffff;
        const int at = valid ? q0 + skew * m : TPP_;        // staged index of x[q0] (q0 / D = m); idle lanes read inside the block
        // this phase's 151 taps: one coalesced read per warp, then broadcast reads from shared memory
        __syncwarp();
        for (int i2 = lane; i2 < HPITCH; i2 += 32) wt[i2] = __ldg(hp + (long long)ph * HPITCH + i2);
        __syncwarp();
        float acc = 0.0f;
        const float *xp = xs_sh + at;
        if (__all_sync(0xffffffffu, q0 >= TPP_ - 1)) {      // every tap of every lane is inside the block
            if (!SKEW) {
#pragma unroll 2
                for (int c = 0; c + 4 <= TPP_; c += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(wt + c);
                    const float v0 = xp[-c], v1 = xp[-c - 1], v2 = xp[-c - 2], v3 = xp[-c - 3];
                    acc = EXACT ? __fadd_rn(acc, __fmul_rn(v0, t.x)) : fmaf(v0, t.x, acc);
                    acc = EXACT ? __fadd_rn(acc, __fmul_rn(v1, t.y)) : fmaf(v1, t.y, acc);
                    acc = EXACT ? __fadd_rn(acc, __fmul_rn(v2, t.z)) : fmaf(v2, t.z, acc);
                    acc = EXACT ? __fadd_rn(acc, __fmul_rn(v3, t.w)) : fmaf(v3, t.w, acc);
                }
#pragma unroll
                for (int c = TPP_ / 4 * 4; c < TPP_; ++c) acc = EXACT ? __fadd_rn(acc, __fmul_rn(xp[-c], wt[c])) : fmaf(xp[-c], wt[c], acc);
            } else {
                const int cx = min(qr + 1, TPP_);           // taps c <= qr stay in row m, the rest sit one pad word further down
#pragma unroll 4
                for (int c = 0; c < cx; ++c) acc = EXACT ? __fadd_rn(acc, __fmul_rn(xp[-c], wt[c])) : fmaf(xp[-c], wt[c], acc);
#pragma unroll 4
                for (int c = cx; c < TPP_; ++c) acc = EXACT ? __fadd_rn(acc, __fmul_rn(xp[-c - 1], wt[c])) : fmaf(xp[-c - 1], wt[c], acc);
            }
        } else {                                            // some lane reaches into the history: branch-free select per tap
            auto tap = [&](int c, float t, float hv) {
                const int d = q0 - c;                       // >= 0: inside the block
                const int idx = d >= 0 ? d + (SKEW ? m - (c > qr ? 1 : 0) : 0) : 0;
                const float xv = xs_sh[idx];
                const float v = d >= 0 ? xv : hv;
                acc = EXACT ? __fadd_rn(acc, __fmul_rn(v, t)) : fmaf(v, t, acc);
            };
#pragma unroll 2
            for (int c = 0; c + 4 <= TPP_; c += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(wt + c), hv = *reinterpret_cast<const float4 *>(hist + c);
                tap(c, t.x, hv.x); tap(c + 1, t.y, hv.y); tap(c + 2, t.z, hv.z); tap(c + 3, t.w, hv.w);
            }
#pragma unroll
            for (int c = TPP_ / 4 * 4; c < TPP_; ++c) tap(c, wt[c], hist[c]);
        }
        if (a.gain_up) acc = __fmul_rn(acc, (float)U);
        if (valid) ys[o] = acc;
    }
}

__global__ void resample_state_kernel(const ResDev a, int i0, int i1) {
    const int s = blockIdx.y;
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const int p = a.n_ref - a.nzi - 1 + i;
    if (p < 0 || p >= a.n) return;
    a.zi[(long long)s * a.nzi + i] = a.x[(long long)s * a.ldx + (long long)(a.n_blocks - 1) * a.n + p];
}

// the same copy four samples at a time, for geometries where source, destination and length are all 16-byte multiples
// (the RDS resampler: 2868 samples from offset 12492)
__global__ void resample_state4_kernel(const ResDev a) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * i >= a.nzi) return;
    const float4 *src = reinterpret_cast<const float4 *>(a.x + (long long)s * a.ldx + (long long)(a.n_blocks - 1) * a.n + (a.n_ref - a.nzi - 1));
    reinterpret_cast<float4 *>(a.zi + (long long)s * a.nzi)[i] = __ldg(src + i);
}

// ---------------------------------------------------------------------------------------------------------------
// RDS decoder: one stream per lane
// ---------------------------------------------------------------------------------------------------------------
constexpr int SPS = 24;  // samples per chip at 57 kHz
// parity-check matrix rows as 10-bit words (MSB = syndrome element 0), src/fm_radio.cpp:477; offsets A-D, :479-482
__constant__ uint16_t kH[26] = {0x200, 0x100, 0x080, 0x040, 0x020, 0x010, 0x008, 0x004, 0x002, 0x001, 0x2DC, 0x16E, 0x0B7,
                                0x287, 0x39F, 0x313, 0x355, 0x376, 0x1BB, 0x201, 0x3DC, 0x1EE, 0x0F7, 0x2A7, 0x38F, 0x31B};
__constant__ uint16_t kSyn[4] = {0x3D8, 0x3D4, 0x25C, 0x258};
// the same matrix by columns: bit j of kHcol[b] = bit b of kH[j], so that bit b of the syndrome of a 26-bit window w (bit j = the
// window's j-th bit) is the parity of w & kHcol[b] -- ten AND + POPC on registers instead of 26 loads from a byte array per position
__constant__ uint32_t kHcol[10] = {0x3cdf200, 0x3e6f900, 0x1f37c80, 0x3344c40, 0x257d420, 0xe61810, 0x730c08, 0x1f47404, 0x337c802, 0x39be401};

enum { W_BLOCK = 0, W_OFFSET, W_START, W_LONELY, W_FRONT, W_PREBIT, W_NBITS, W_PRINTPOS, W_LASTPOS1, W_BAD, W_CARRY = 10, W_BITS = 40 };

__device__ __forceinline__ bool same_sign(float a, float b) { return (a > 0 && b > 0) || (a < 0 && b < 0); }

// One station per lane, 32 stations per (single-warp) CTA.  Every global access is made by the warp TOGETHER -- a station's 160
// state words, its 152 symbols (stride 24 in the RRC buffer), its 24 first samples in block 0, its decoded bits -- and staged in
// shared memory with an odd row stride, so the lanes then walk their own rows without bank conflicts.  (First version: each lane
// loaded its own state and symbols, ~300 dependent global loads of one sector each per block: 0.10 ms at 1.7 % warps active.)
constexpr int DEC_ROW = FMRX_RDS_STATE_WORDS + 1;  // 161: odd stride
constexpr int DEC_BITS_ROW = 21;                   // 80 bytes of bits + one pad word

__device__ __forceinline__ unsigned window_bit(unsigned long long lo, unsigned long long hi, int g) {
    return (unsigned)(((g < 64) ? (lo >> g) : (hi >> (g - 64))) & 1ull);
}

__global__ void __launch_bounds__(32) rds_decode_kernel(const float *rrc, long long ld, int n_streams, int n_blocks, int n, uint8_t *bits_out, int32_t *n_bits_out,
                                                        fmrx_rds_event *events, int32_t *n_events_out, int32_t *state) {
    __shared__ int32_t sst[32 * DEC_ROW];
    __shared__ float ssym[32 * DEC_ROW];
    __shared__ float sdense[32 * (SPS + 1)];
    __shared__ uint32_t sbits[32 * DEC_BITS_ROW];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x, s0 = blockIdx.x * 32, s = s0 + lane;
    const int live = min(32, n_streams - s0);
    const bool mine = lane < live;
    // eight stations' loads are issued before the first of them is stored: a single warp per SM has nothing else to hide a DRAM
    // round trip behind (one station at a time -- load, store, next -- took longer than the per-lane loads it replaced)
    static_assert(FMRX_RDS_STATE_WORDS == 5 * 32, "five state words per lane per station");
#pragma unroll 1
    for (int t0 = 0; t0 < 32; t0 += 8) {
        int32_t v[8][5];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int m = 0; m < 5; ++m) v[u][m] = t0 + u < live ? state[(long long)(s0 + t0 + u) * FMRX_RDS_STATE_WORDS + lane + 32 * m] : 0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int m = 0; m < 5; ++m) sst[(t0 + u) * DEC_ROW + lane + 32 * m] = v[u][m];
    }
    __syncwarp();
    int32_t *st = sst + lane * DEC_ROW;            // bits[i] = st[W_BITS + i]: `bit_stream`, persistent (Q12)
    uint8_t *mybits = reinterpret_cast<uint8_t *>(sbits + lane * DEC_BITS_ROW);
    int block_id = mine ? st[W_BLOCK] : 1;
    unsigned offset = mine ? (unsigned)st[W_OFFSET] : 0u, start_pos = mine ? (unsigned)st[W_START] : 0u;
    float lonely = __int_as_float(st[W_LONELY]);
    int front_bit = st[W_FRONT], prebit = st[W_PREBIT], nbits = st[W_NBITS];
    unsigned printpos = (unsigned)st[W_PRINTPOS];
    int last_pos = st[W_LASTPOS1] - 1, bad = st[W_BAD];
    const int nsym = n / SPS;

    for (int b = 0; b < n_blocks; ++b, ++block_id) {
        const float *r0 = rrc + (long long)s0 * ld + (long long)b * n;  // station t of this CTA: r0 + t * ld
        if (__any_sync(FULL, mine && block_id == 0)) {  // :503-517, the sampling phase from |rrc[0..23]| of a station's first block
            for (int t = 0; t < live; ++t)
                if (lane < SPS) sdense[t * (SPS + 1) + lane] = r0[(long long)t * ld + lane];
            __syncwarp();
            if (mine && block_id == 0) {
                const float *d = sdense + lane * (SPS + 1);
                float best = fabsf(d[0]);
                for (unsigned i = 1; i < SPS; ++i)
                    if (fabsf(d[i]) > best) { best = fabsf(d[i]); offset = i; }
            }
        }
#pragma unroll 1
        for (int t0 = 0; t0 < 32; t0 += 8) {  // sym[k] = r[24k + offset] of every station of the CTA, eight stations in flight
            float v[8][5];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned off_t = __shfl_sync(FULL, offset, t0 + u);
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    const int k = lane + 32 * m;
                    v[u][m] = (t0 + u < live && k < nsym) ? r0[(long long)(t0 + u) * ld + SPS * k + off_t] : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int m = 0; m < 5; ++m) ssym[(t0 + u) * DEC_ROW + lane + 32 * m] = v[u][m];
        }
        __syncwarp();
        int nd = 0, nev = 0;
        if (mine) {
            const float *sym = ssym + lane * DEC_ROW;
            if (block_id == 0) {  // :542-558 (loop index starts at 0, Q10)
                int c0 = 0, c1 = 0;
                for (int j = 0; j < nsym / 4; ++j) {
                    const float a0 = sym[2 * j], a1 = sym[2 * j + 1], a2 = sym[2 * j + 2];
                    if (same_sign(a0, a1)) ++c0;
                    else if (same_sign(a1, a2)) ++c1;
                }
                if (c0 > c1) start_pos = 1;
                else if (c1 > c0) start_pos = 0;
            }
            const int want = nsym / 2 - (int)start_pos;  // :560
            for (int i = nbits; i < want; ++i) st[W_BITS + i] = 0;
            nbits = want;
            if (start_pos == 1 && block_id != 0) {  // :565-572
                const float f0 = sym[0];
                if (lonely > f0) front_bit = 1;
                else if (f0 > lonely) front_bit = 0;
            }
            for (int k = 0; k < nbits; ++k) {  // :574-585
                const unsigned a = 2u * k + start_pos;
                if (a + 1 > (unsigned)nsym - 1) break;
                const float u = sym[a], v = sym[a + 1];
                if (u > v) st[W_BITS + k] = 1;
                else if (u < v) st[W_BITS + k] = 0;
            }
            if (start_pos == 1) {  // :587-592
                for (int i = nbits; i > 0; --i) st[W_BITS + i] = st[W_BITS + i - 1];
                st[W_BITS] = front_bit;
                ++nbits;
                lonely = sym[nsym - 1];
            }
            int off = 0;  // :596-616
            if (block_id == 0) { prebit = st[W_BITS]; off = 1; }
            const int ncarry = block_id != 0 ? 27 : 0;
            // the (at most 27 carried + 77 new) differentially decoded bits packed into two words; the window at `pos` is bits pos .. pos + 25
            unsigned long long lo = 0ull, hi = 0ull;
            for (int g = 0; g < ncarry; ++g) lo |= (unsigned long long)(st[W_CARRY + g] & 1) << g;
            nd = nbits - off;
            for (int t = 0; t < nd; ++t) {
                const unsigned v = (unsigned)(prebit ^ st[W_BITS + t + off]) & 1u;
                const int g = ncarry + t;
                if (g < 64) lo |= (unsigned long long)v << g;
                else hi |= (unsigned long long)v << (g - 64);
                prebit = st[W_BITS + t + off];
                mybits[t] = (uint8_t)v;
            }
            prebit = st[W_BITS + nbits - 1];
            const int total = ncarry + nd;
            fmrx_rds_event *ev = events ? events + ((long long)s * n_blocks + b) * FMRX_MAX_EVENTS : nullptr;
            unsigned pos = 0;
            for (;;) {  // :631-713
                const unsigned long long sh = pos == 0 ? lo : pos < 64 ? ((lo >> pos) | (hi << (64 - pos))) : (hi >> (pos - 64));
                const unsigned w = (unsigned)sh & 0x3FFFFFFu;
                unsigned syn = 0;
#pragma unroll
                for (int bb = 0; bb < 10; ++bb) syn |= (unsigned)(__popc(w & kHcol[bb]) & 1) << bb;
                for (int L = 0; L < 4; ++L) {
                    if (syn != kSyn[L]) continue;
                    const bool good = last_pos == -1 || printpos - (unsigned)last_pos == 26u;
                    if (ev && nev < FMRX_MAX_EVENTS) { ev[nev].block = block_id; ev[nev].kind = good ? FMRX_EV_GOOD : FMRX_EV_FALSE; ev[nev].letter = L; ev[nev].position = printpos; }
                    ++nev;
                    if (good) { last_pos = (int)printpos; bad = 0; }
                    else ++bad;
                    break;
                }
                if (bad > 10) {
                    if (ev && nev < FMRX_MAX_EVENTS) { ev[nev].block = block_id; ev[nev].kind = FMRX_EV_RESYNC; ev[nev].letter = -1; ev[nev].position = printpos; }
                    ++nev;
                    bad = 0;
                    last_pos = -1;
                }
                pos += 1;
                if (pos + 26 > (unsigned)total - 1) break;
                printpos += 1;
            }
            for (int g = 0; g < 27; ++g) st[W_CARRY + g] = (int)window_bit(lo, hi, (int)pos - 1 + g);  // :715-718
            if (n_bits_out) n_bits_out[(long long)s * n_blocks + b] = nd;
            if (n_events_out) n_events_out[(long long)s * n_blocks + b] = nev < FMRX_MAX_EVENTS ? nev : FMRX_MAX_EVENTS;
        }
        __syncwarp();
        if (bits_out)
            for (int t = 0; t < live; ++t) {
                const int nd_t = __shfl_sync(FULL, nd, t);
                const uint8_t *src = reinterpret_cast<const uint8_t *>(sbits + t * DEC_BITS_ROW);
                uint8_t *bo = bits_out + ((long long)(s0 + t) * n_blocks + b) * FMRX_MAX_BITS;
                for (int i = lane; i < nd_t; i += 32) bo[i] = src[i];
            }
        __syncwarp();
    }
    if (mine) {
        st[W_BLOCK] = block_id; st[W_OFFSET] = (int)offset; st[W_START] = (int)start_pos; st[W_LONELY] = __float_as_int(lonely);
        st[W_FRONT] = front_bit; st[W_PREBIT] = prebit; st[W_NBITS] = nbits; st[W_PRINTPOS] = (int)printpos;
        st[W_LASTPOS1] = last_pos + 1; st[W_BAD] = bad;
    }
    __syncwarp();
    for (int t = 0; t < live; ++t)
        for (int w = lane; w < FMRX_RDS_STATE_WORDS; w += 32) state[(long long)(s0 + t) * FMRX_RDS_STATE_WORDS + w] = sst[t * DEC_ROW + w];
}

}  // namespace

int launch_resample(const ResampleJob &j, fmrx_stream_t st) {
    ResDev d;
    const int fast = launch_resample_tiled(j, st);  // the RDS 19/80 geometry; anything else falls through to the general kernel
    if (fast > 0) return fast;
    d.x = j.x; d.y = j.y; d.zi = j.zi; d.h = j.h; d.ldx = j.ldx; d.ldy = j.ldy;
    d.n = j.n; d.n_ref = j.n_ref;
    d.ny = j.ny; d.n_blocks = j.n_blocks; d.ntaps = j.ntaps; d.nzi = j.nzi; d.decim = j.decim; d.up = j.up; d.gain_up = j.gain_up;
    dim3 grid((j.ny + 127) / 128, j.n_blocks, j.n_streams);
    // phase-grouped kernel: needs the phase-major tap table, 151 taps per phase, a whole block in shared memory and a lane
    // stride (D, or D + 1 with the pad) that is odd; an even D below 152 would cross the pad twice inside one output
    const int skew = (j.decim % 2 == 0) ? 1 : 0;
    const size_t smem = ((size_t)j.n + (skew ? j.n / j.decim + 1 : 0)) * sizeof(float);
    static const bool no_phase = [] { const char *v = getenv("FMRX_RESAMPLE"); return v && atoi(v) == 1; }();
    if (fast < 0 && !no_phase && j.hp && j.ntaps == kTaps * j.up && j.n_ref == j.n && j.n % 4 == 0 && j.ldx % 4 == 0 && ((uintptr_t)j.x & 15) == 0 &&
        (!skew || j.decim >= kTaps + 1) && j.nzi >= kTaps && j.n_ref >= j.nzi + 1 && j.n > kTaps && smem <= 200 * 1024) {
        auto kern = j.exact ? (skew ? resample_phase_kernel<true, true> : resample_phase_kernel<true, false>)
                            : (skew ? resample_phase_kernel<false, true> : resample_phase_kernel<false, false>);
        cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ea) return (int)ea;
        // outputs whose taps reach into the history: o < ceil(150 * U / D).  When they are a handful (29 of 2949 for x24 /125)
        // they go to the general kernel so that every task of the phase kernel runs its all-in-block loop; the Q14 call
        // (x24 /5: 720 of 2949) keeps them in the phase kernel's select loop, and so does x147 /800 (28 outputs, but the
        // general kernel gathers from an 89 KB tap table there: 0.2 ms for them against 0.05 ms in the select loop)
        int o_edge = (int)(((long long)(kTaps - 1) * j.up + j.decim - 1) / j.decim);
        if (o_edge > j.ny) o_edge = j.ny;
        const int o_skip = (!skew && o_edge <= 64) ? o_edge : 0;
        kern<<<dim3(j.n_blocks, j.n_streams), 256, smem, st>>>(d, j.hp, PairConst{-0.0f, 1.0f}, o_skip);
        launch_counter() += 1;
        if (o_skip > 0) {
            dim3 ge(1, j.n_blocks, j.n_streams);
            if (j.exact) resample_kernel<true><<<ge, 64, 0, st>>>(d, o_skip);
            else resample_kernel<false><<<ge, 64, 0, st>>>(d, o_skip);
            launch_counter() += 1;
        }
    } else if (fast < 0) {
        if (j.exact) resample_kernel<true><<<grid, 128, 0, st>>>(d, j.ny);
        else resample_kernel<false><<<grid, 128, 0, st>>>(d, j.ny);
        launch_counter() += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e) return (int)e;
    const int p0 = j.n_ref - j.nzi - 1;
    const bool vec = (j.nzi & 3) == 0 && (p0 & 3) == 0 && p0 >= 0 && p0 + j.nzi <= j.n && (j.ldx & 3) == 0 && (j.n & 3) == 0 &&
                     (((uintptr_t)j.x | (uintptr_t)j.zi) & 15) == 0;
    if (j.live_state_only && j.nzi > kTaps) {  // entries (nzi-1-c)/up, c = 0..150
        const int i0 = (j.nzi - kTaps) / j.up, i1 = (j.nzi - 1) / j.up + 1;
        dim3 sg((i1 - i0 + 255) / 256, j.n_streams);
        resample_state_kernel<<<sg, 256, 0, st>>>(d, i0, i1);
    } else if (vec) {
        dim3 sg((j.nzi / 4 + 255) / 256, j.n_streams);
        resample_state4_kernel<<<sg, 256, 0, st>>>(d);
    } else {
        dim3 sg((j.nzi + 255) / 256, j.n_streams);
        resample_state_kernel<<<sg, 256, 0, st>>>(d, 0, j.nzi);
    }
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

int launch_rds_decode(const float *rrc, long long ld, int n_streams, int n_blocks, int n, uint8_t *bits, int32_t *n_bits,
                      fmrx_rds_event *events, int32_t *n_events, int32_t *state, fmrx_stream_t st) {
    rds_decode_kernel<<<(n_streams + 31) / 32, 32, 0, st>>>(rrc, ld, n_streams, n_blocks, n, bits, n_bits, events, n_events, state);
    launch_counter() += 1;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
