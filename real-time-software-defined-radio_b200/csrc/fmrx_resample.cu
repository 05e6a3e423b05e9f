// Rational polyphase resampler, register-tiled fast path for the geometry the receive chain runs at full rate:
// the RDS 19/80 resampler with its 151-taps-per-phase anti-image filter (SURVEY 8a row a12;
// /root/reference/src/filter.cpp:301-339 convolveWithDecimMode1RDS, call site src/fm_radio.cpp:408).
//
// Reference index map (App. D.2): output o visits the taps k = k0 + U*c, c = 0..150, with k0 = (D*o) mod U, and reads
// x[q0 - c] with q0 = floor(D*o/U) while that index is inside the block; the taps beyond read zi[(Z-1-c)/U] (Q6: an
// index that depends on the tap COUNT, not on the position).  Only the first ceil(150*U/D) outputs of a block have such
// taps, and they come last in the summation order, so the work splits cleanly:
//
//   resample_tile_kernel   every output's in-block sum.  Outputs o and o+U use the same U-th of the taps and windows
//                          D samples apart, so one THREAD owns one group of U consecutive outputs (all U phases) whose
//                          windows lie inside 226 consecutive samples: each staged sample is read once (128-bit LDS
//                          from a tile with a row pitch of D+4 words: 16-byte aligned and conflict-free) and applied
//                          to every phase it feeds with the tap as a constant-bank operand (the table arrives by
//                          value in the parameter block), newest sample first = taps ascending = the reference's
//                          summation order.  The unrolled body of all 19 phases would be 46 KB of FFMAs, more than the
//                          instruction cache holds, so the phases are cut into three launches of <= 7 phases (<= 20 KB).
//   resample_fix_kernel    the history taps of the first outputs of every block, continued from the stored partial sum
//                          in the reference's order, then the x U gain (src/filter.cpp:333).
//
// Every other geometry (mode 1's 24/125 with its NaN tap, the 147/800 function-level case, the exact two-rounding
// variant) stays on the one-thread-per-output kernel in fmrx_rds.cu.
#include <cuda_runtime.h>

#include "fmrx_internal.h"

namespace fmrx {
namespace {

constexpr int TPP = kTaps;  // taps per phase

// Taps regrouped for the walk: row r holds phase r's taps h[(D*r)%U + U*c] at element ROW_OFF(r) + c, where the offset
// makes the four taps a quad of samples needs (c, c+1, c+2, c+3 for the newest..oldest sample of the quad) one
// 16-byte-aligned group, so they reach the uniform registers as one 128-bit constant load per four FFMAs.
constexpr int ROWQ = 41;  // float4 per row: 4 + 3 + 151 + 3 rounded up
template <int U>
struct PolyTaps {
    float4 q[U][ROWQ];
};

template <int U, int D>
struct PolyGeom {
    static_assert(D % 4 == 0, "rows must keep 128-bit alignment");
    static constexpr int NT = 64;                      // groups (threads) per CTA
    static constexpr int LEAD = 152;                   // staged samples ahead of the first group's origin: >= 150, multiple of 4
    static constexpr int QMAX = (D * (U - 1)) / U;     // newest sample (relative to the group origin) any phase reads
    static constexpr int WIN = LEAD + QMAX + 1;        // samples one thread touches
    static constexpr int WIN4 = (WIN + 3) / 4;
    static constexpr int SPAN = D * (NT - 1) + 4 * WIN4;
    static constexpr int PITCH = D + 4;
    static constexpr int WORDS = ((SPAN + D - 1) / D) * PITCH;
    static constexpr int NFIX = (150 * U + D - 1) / D;  // outputs o with floor(D*o/U) < 150 have history taps
    __host__ __device__ static constexpr int phys(int i) { return i + 4 * (i / D); }
    __host__ __device__ static constexpr int q0(int r) { return (D * r) / U; }
    __host__ __device__ static constexpr int row_off(int r) { return 4 + (((3 - q0(r)) % 4) + 4) % 4; }
};

struct PolyDev {
    const float *x;
    float *y;
    const float *zi;
    const float *h;  // device copy of the taps (fix-up kernel)
    long long ldx, ldy;
    int n, n_ref, ny, nzi, n_blocks, gain_up;
};

template <int U, int D, int R0, int R1>
__global__ void __launch_bounds__(64) resample_tile_kernel(const PolyDev a, const __grid_constant__ PolyTaps<U> taps) {
    using G = PolyGeom<U, D>;
    __shared__ __align__(16) float sm[G::WORDS];
    const int s = blockIdx.z, b = blockIdx.y, g0 = blockIdx.x * G::NT;
    const float *xb = a.x + (long long)s * a.ldx + (long long)b * a.n;
    const int P0 = D * g0 - G::LEAD;  // block-relative position of staged sample 0 (a multiple of 4)
    if ((a.n & 3) == 0 && ((uintptr_t)xb & 15) == 0) {
        // asynchronous 16-byte copies straight into the padded tile (LDGSTS): every quad of a thread is in flight at once
        // and none of them occupies a register.  P0 and n are multiples of 4, so a quad is wholly inside the block or
        // wholly outside it; outside = zero fill (src-size 0): those positions contribute nothing to the in-block sum.
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
        for (int j = threadIdx.x; j < G::SPAN / 4; j += G::NT) {
            const int p = P0 + 4 * j;
            const bool in = p >= 0 && p + 4 <= a.n;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sbase + 4u * (unsigned)G::phys(4 * j)), "l"(in ? xb + p : xb), "r"(in ? 16 : 0));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int j = threadIdx.x; j < G::SPAN / 4; j += G::NT) {
            const int p = P0 + 4 * j;
            float4 v;
            v.x = (p >= 0 && p < a.n) ? xb[p] : 0.0f;
            v.y = (p + 1 >= 0 && p + 1 < a.n) ? xb[p + 1] : 0.0f;
            v.z = (p + 2 >= 0 && p + 2 < a.n) ? xb[p + 2] : 0.0f;
            v.w = (p + 3 >= 0 && p + 3 < a.n) ? xb[p + 3] : 0.0f;
            *reinterpret_cast<float4 *>(sm + G::phys(4 * j)) = v;
        }
    }
    __syncthreads();

    const float *w = sm + G::PITCH * threadIdx.x;  // logical index D*t + j  ->  PITCH*t + phys(j), j < D*? handled by phys
    float acc[R1 - R0];
#pragma unroll
    for (int r = 0; r < R1 - R0; ++r) acc[r] = 0.0f;
#pragma unroll
    for (int q = G::WIN4 - 1; q >= 0; --q) {
        // skip the quads no phase of this launch reads (compile time)
        bool used = false;
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int r = R0; r < R1; ++r) {
                const int c = (D * r) / U - (4 * q + e - G::LEAD);
                used = used || (c >= 0 && c < TPP);
            }
        if (!used) continue;
        const float4 v = *reinterpret_cast<const float4 *>(w + G::phys(4 * q));
#pragma unroll
        for (int r = R0; r < R1; ++r) {
            const int c0 = G::q0(r) - (4 * q + 3 - G::LEAD);  // tap count of the quad's newest sample; c0+1.. for the older ones
            if (c0 + 3 < 0 || c0 >= TPP) continue;
            const float4 t = taps.q[r][(G::row_off(r) + c0) / 4];
            float &ac = acc[r - R0];
            if (c0 >= 0 && c0 < TPP) ac = fmaf(v.w, t.x, ac);
            if (c0 + 1 >= 0 && c0 + 1 < TPP) ac = fmaf(v.z, t.y, ac);
            if (c0 + 2 >= 0 && c0 + 2 < TPP) ac = fmaf(v.y, t.z, ac);
            if (c0 + 3 >= 0 && c0 + 3 < TPP) ac = fmaf(v.x, t.w, ac);
        }
    }
    const int g = g0 + threadIdx.x;
    const int o = U * g + R0;
    if (U * g >= a.ny) return;
    float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny + o;
    const float gain = a.gain_up ? (float)U : 1.0f;
#pragma unroll
    for (int r = 0; r < R1 - R0; ++r)
        if (o + r < a.ny) ys[r] = (o + r < G::NFIX) ? acc[r] : __fmul_rn(acc[r], gain);  // the first NFIX are finished by the fix-up kernel
}

// one warp per (stream, block): lanes over the outputs that have history taps
template <int U, int D>
__global__ void resample_fix_kernel(const PolyDev a, int total) {
    using G = PolyGeom<U, D>;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= total) return;
    const int s = wid / a.n_blocks, b = wid % a.n_blocks;
    const float *xs = a.x + (long long)s * a.ldx;
    const float *zs = a.zi + (long long)s * a.nzi;
    float *ys = a.y + (long long)s * a.ldy + (long long)b * a.ny;
    // history values live at indices (nzi-1-c)/U of the carried state (block 0) or of the previous block's tail
    const float *hist = b > 0 ? xs + (long long)(b - 1) * a.n + (a.n_ref - a.nzi - 1) : zs;
    for (int o = lane; o < G::NFIX && o < a.ny; o += 32) {
        const int base = D * o, k0 = base % U, q0 = base / U;
        const float *hp = a.h + k0;
        float acc = ys[o];
        int c = q0 + 1;
        for (; c + 8 <= TPP; c += 8) {  // loads of a batch are independent of the sum: all in flight before the eight dependent FMAs
            float t[8], v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { t[u] = __ldg(hp + U * (c + u)); v[u] = __ldg(hist + (a.nzi - 1 - (c + u)) / U); }
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fmaf(v[u], t[u], acc);
        }
        for (; c < TPP; ++c) acc = fmaf(__ldg(hist + (a.nzi - 1 - c) / U), __ldg(hp + U * c), acc);
        ys[o] = a.gain_up ? __fmul_rn(acc, (float)U) : acc;
    }
}

}  // namespace

// returns -1 when the job is not the fast path's geometry (the caller then uses the general kernel)
int launch_resample_tiled(const ResampleJob &j, fmrx_stream_t st) {
    constexpr int U = 19, D = 80;
    using G = PolyGeom<U, D>;
    if (j.up != U || j.decim != D || j.ntaps != TPP * U || j.nzi != TPP * U - 1 || j.exact || !j.h_host || j.ny % U != 0 || j.ny > j.n_ref * U / D ||
        j.n_ref - j.nzi - 1 < 0 || j.n < G::NFIX)
        return -1;
    PolyDev d;
    d.x = j.x; d.y = j.y; d.zi = j.zi; d.h = j.h; d.ldx = j.ldx; d.ldy = j.ldy;
    d.n = j.n; d.n_ref = j.n_ref; d.ny = j.ny; d.nzi = j.nzi; d.n_blocks = j.n_blocks; d.gain_up = j.gain_up;
    static_assert((G::row_off(0) + G::q0(0) - 3) % 4 == 0 && G::row_off(U - 1) + TPP + 3 <= 4 * ROWQ, "tap row layout");
    PolyTaps<U> t;
    float *tf = reinterpret_cast<float *>(t.q);
    for (int i = 0; i < U * ROWQ * 4; ++i) tf[i] = 0.0f;
    for (int r = 0; r < U; ++r)
        for (int c = 0; c < TPP; ++c) tf[r * ROWQ * 4 + G::row_off(r) + c] = j.h_host[(D * r) % U + U * c];
    const int groups = j.ny / U;
    dim3 grid((groups + G::NT - 1) / G::NT, j.n_blocks, j.n_streams);
    resample_tile_kernel<U, D, 0, 7><<<grid, G::NT, 0, st>>>(d, t);
    resample_tile_kernel<U, D, 7, 13><<<grid, G::NT, 0, st>>>(d, t);
    resample_tile_kernel<U, D, 13, 19><<<grid, G::NT, 0, st>>>(d, t);
    const int total = j.n_streams * j.n_blocks;
    resample_fix_kernel<U, D><<<(total * 32 + 127) / 128, 128, 0, st>>>(d, total);
    launch_counter() += 4;
    return (int)cudaGetLastError();
}

}  // namespace fmrx
