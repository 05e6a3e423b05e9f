// FP32 issue-rate microbenchmarks: the roofline denominators for the FIR kernels.  The FIR stages are FP32-pipe bound
// (25-114 flop per input byte, far right of the HBM ridge), and MEASURED_PEAKS.json only carries HBM and
// bf16-tensor peaks, so the FP32 peak is measured here, on the same device, by the same binary.
//   kind 0: FFMA            (1 lane-op = 1 fused multiply-add)
//   kind 1: FMUL + FADD     (the reference-exact tap: two roundings, two lane-ops)
//   kind 2: FFMA2           (sm_100 packed fp32x2: two FMAs per lane per instruction)
//   kind 3: FMUL2 + FADD2
// and the issue rates the PLL kernel's bound is made of (its step is ~1/3 FP64 and ~1/8 conversions, csrc/fmrx_pllmath.h):
//   kind 4: DFMA            (FP64 pipe)
//   kind 5: F2F.F64.F32 + F2F.F32.F64 pairs (the conversion pipe; 1 lane-op = one conversion)
//   kind 6: SHF + LOP3      (integer ALU pipe)
//   kind 7: 4 DFMA + one conversion pair, interleaved (1 lane-op = one DFMA or one conversion): do the two pipes overlap?
//   kind 8: DFMA with three distinct vector-register operands (kind 4's multiplier and addend are kernel parameters, i.e. uniform)
// Reported as tera lane-ops per second.
#include <cuda_runtime.h>

#include <cstdlib>

#include "fmrx_internal.h"
#include "fmrx_pllmath.h"

namespace fmrx {
namespace {

// The PLL kernel's roofline is not a throughput: one loop is one dependency chain (detector -> loop filter -> oscillator
// -> next sample), 8192 loops are 256 warps on 592 schedulers, and what bounds the kernel is the LATENCY of one step of
// that chain.  This measures it in isolation: one warp, the step of csrc/fmrx_pllmath.h fed from registers (no memory
// traffic, no other warp on the scheduler), cycles per step from clock64.
template <int V>
__global__ void pll_chain_kernel(float *out, long long *cycles, int steps, float freq_ratio, float scale) {
    using namespace pllmath;
    __shared__ __align__(16) PllTheta theta[16];
    __shared__ double kdoubles[kPllKDoubles];
    if (threadIdx.x < 16) theta[threadIdx.x] = pll_theta_entry(threadIdx.x);
    if (threadIdx.x < kPllKDoubles) kdoubles[threadIdx.x] = pll_k_value(threadIdx.x);
    __syncwarp();
    const PllK K = pll_k_from(kdoubles, 0x38000000u);
    PllCarry c{0.0f, 0.0f, 1.0f, 0.0f};
    PllFast f;
    pll_disarm(f);
    const PllCoef p{1e-6f * 3.555f, 1e-3f * 2.666f, scale, 0.1f, (2 * 3.14159265358979323846) * (double)freq_ratio};
    float acc = 0.0f, x = 0.05f + 1e-4f * threadIdx.x;
    { const PllLibmOut o = pll_step_libm(c, p, x, 1.0f); c = o.c; if (V >= 1) pll_rearm1(f, o.trig, true, theta, K); else pll_rearm(f, o.trig); }
    int bad = 0;
    const long long t0 = clock64();
    for (int k = 1; k < steps; k += 4) {
        bool ok0, ok1, ok2, ok3;
        const float x0 = -x * 0.999f + 1e-5f, x1 = -x0 * 0.999f + 1e-5f, x2 = -x1 * 0.999f + 1e-5f, x3 = -x2 * 0.999f + 1e-5f;
        x = x3;
        if (V >= 1) {  // the input alternates in sign: x0 < 0 < x1 ...
            acc += pll_step_fast1<V == 2>(c, f, p, K, x0, __fadd_rn((float)k, 1.0f), x1 < 0.0f, theta, ok0);
            acc += pll_step_fast1<V == 2>(c, f, p, K, x1, __fadd_rn((float)k, 2.0f), x2 < 0.0f, theta, ok1);
            acc += pll_step_fast1<V == 2>(c, f, p, K, x2, __fadd_rn((float)k, 3.0f), x3 < 0.0f, theta, ok2);
            acc += pll_step_fast1<V == 2>(c, f, p, K, x3, __fadd_rn((float)k, 4.0f), x3 > 0.0f, theta, ok3);
        } else {
            acc += pll_step_fast(c, f, p, x0, __fadd_rn((float)k, 1.0f), ok0);
            acc += pll_step_fast(c, f, p, x1, __fadd_rn((float)k, 2.0f), ok1);
            acc += pll_step_fast(c, f, p, x2, __fadd_rn((float)k, 3.0f), ok2);
            acc += pll_step_fast(c, f, p, x3, __fadd_rn((float)k, 4.0f), ok3);
        }
        bad += !(ok0 && ok1 && ok2 && ok3);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = acc + c.phase + bad;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

constexpr int ILP = 16;
constexpr int ITERS = 4096;

template <int KIND>
__global__ void __launch_bounds__(256) fp32_rate_kernel(float *sink, float a, float b) {
    float2 acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-3f - i);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) { acc[i].x = fmaf(acc[i].x, a, b); acc[i].y = fmaf(acc[i].y, a, b); }
            if (KIND == 1) { acc[i].x = __fadd_rn(__fmul_rn(acc[i].x, a), b); acc[i].y = __fadd_rn(__fmul_rn(acc[i].y, a), b); }
            if (KIND == 2) acc[i] = __ffma2_rn(acc[i], a2, b2);
            if (KIND == 3) acc[i] = __fadd2_rn(__fmul2_rn(acc[i], a2), b2);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) sink[0] = s;  // keep the chain alive without a store in practice
}

template <int KIND>
__global__ void __launch_bounds__(256) pipe_rate_kernel(float *sink, double a, double b, int iters) {
    double d[ILP];
    float f[ILP];
    unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x * 2654435761u + i; }
    const unsigned k1 = (unsigned)__double2loint(a) | 1u, k2 = (unsigned)__double2hiint(b);
    double e[ILP], g[ILP];  // kind 8: per-thread values, so they live in vector registers
#pragma unroll
    for (int i = 0; i < ILP; ++i) { e[i] = a + 1e-9 * (threadIdx.x + i); g[i] = b + 1e-12 * (threadIdx.x * 3 + i); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 4) d[i] = __fma_rn(d[i], a, b);
            if (KIND == 5) {  // F2F.F64.F32 then F2F.F32.F64, kept by the volatile asm (the compiler would fold the round trip away)
                double t;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(t));
            }
            if (KIND == 6) u[i] = __funnelshift_l(u[i], u[i], 7) ^ (k1 + k2);  // SHF + LOP3
            if (KIND == 8) d[i] = __fma_rn(d[i], e[i], g[i]);
            if (KIND == 7) {
                double t;
                d[i] = __fma_rn(d[i], a, b);
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i]));
                d[i] = __fma_rn(d[i], a, b);
                d[i] = __fma_rn(d[i], a, b);
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(t));
                d[i] = __fma_rn(d[i], a, b);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i] + (double)f[i] + (double)u[i];
    if (s == 123.456) sink[0] = (float)s;
}

template <int KIND>
int run_pipe(int reps, double *tera) {
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e) return (int)e;
    float *sink = nullptr;
    e = cudaMalloc(&sink, 4);
    if (e) return (int)e;
    const int grid = prop.multiProcessorCount * 8, iters = 1024;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(t0);
        pipe_rate_kernel<KIND><<<grid, 256>>>(sink, 0.99999, 1e-4, iters);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        // lane-ops per (i, it) per thread: one DFMA; two conversions; two integer instructions
        const double ops = (double)grid * 256 * iters * ILP * ((KIND == 4 || KIND == 8) ? 1.0 : KIND == 7 ? 6.0 : 2.0);
        const double rate = ops / (ms * 1e-3) / 1e12;
        if (r >= 2 && rate > best) best = rate;
    }
    launch_counter() += reps + 2;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    if (e) return (int)e;
    *tera = best;
    return (int)cudaGetLastError();
}

template <int KIND>
int run(int reps, double *tera) {
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e) return (int)e;
    float *sink = nullptr;
    e = cudaMalloc(&sink, 4);
    if (e) return (int)e;
    const int grid = prop.multiProcessorCount * 8;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    for (int r = 0; r < reps + 2; ++r) {
        cudaEventRecord(t0);
        fp32_rate_kernel<KIND><<<grid, 256>>>(sink, 0.9999f, 1e-4f);
        cudaEventRecord(t1);
        e = cudaEventSynchronize(t1);
        if (e) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        // lane-ops: kinds 0/2 = 2 FMAs per (i, it) per thread; kinds 1/3 = 2 muls + 2 adds
        const double ops = (double)grid * 256 * ITERS * ILP * ((KIND == 0 || KIND == 2) ? 2.0 : 4.0);
        const double rate = ops / (ms * 1e-3) / 1e12;
        if (r >= 2 && rate > best) best = rate;
    }
    launch_counter() += reps + 2;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(sink);
    if (e) return (int)e;
    *tera = best;
    return (int)cudaGetLastError();
}

}  // namespace

int measure_pll_chain(double *cycles_per_step) {
    const int steps = 60000;
    float *out = nullptr;
    long long *cyc = nullptr, h = 0;
    cudaError_t e = cudaMalloc(&out, 32 * sizeof(float));
    if (e) return (int)e;
    e = cudaMalloc(&cyc, sizeof(long long));
    if (e) { cudaFree(out); return (int)e; }
    static const int variant = [] { const char *e = getenv("FMRX_PLL_STEP"); return e ? atoi(e) : 2; }();  // the kernel's variant (csrc/fmrx_pll.cu)
    for (int r = 0; r < 2; ++r) {  // second run: instruction cache warm
        if (variant == 0) pll_chain_kernel<0><<<1, 32>>>(out, cyc, steps, 19e3f / 240e3f, 2.0f);
        else if (variant == 2) pll_chain_kernel<2><<<1, 32>>>(out, cyc, steps, 19e3f / 240e3f, 2.0f);
        else pll_chain_kernel<1><<<1, 32>>>(out, cyc, steps, 19e3f / 240e3f, 2.0f);
    }
    launch_counter() += 2;
    e = cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(out); cudaFree(cyc);
    if (e) return (int)e;
    *cycles_per_step = (double)h / (double)(steps - 1);
    return (int)cudaGetLastError();
}

int measure_fp32_peak(int, int kind, int reps, double *tera) {
    if (reps < 1) reps = 1;
    switch (kind) {
        case 0: return run<0>(reps, tera);
        case 1: return run<1>(reps, tera);
        case 2: return run<2>(reps, tera);
        case 3: return run<3>(reps, tera);
        case 4: return run_pipe<4>(reps, tera);
        case 5: return run_pipe<5>(reps, tera);
        case 6: return run_pipe<6>(reps, tera);
        case 7: return run_pipe<7>(reps, tera);
        case 8: return run_pipe<8>(reps, tera);
        default: return (int)cudaErrorInvalidValue;
    }
}

}  // namespace fmrx
