// Double-precision kernels of the PLL recurrence, written for LATENCY: the loop of fmPLL (/root/reference/src/helper.cpp:
// 32-45) is one long dependency chain per stream, atan2 -> loop filter -> sin/cos -> next sample, so what bounds the
// kernel is the depth of that chain, not throughput.  The reference calls the double-precision libm functions on
// float-valued arguments and rounds every result to float; any double result within ~1e-16 of the true value gives the
// same float in all but ~1e-9 of the calls, which is the accuracy these routines keep.
//
//   sincos_cw   Cody-Waite reduction by pi/2 with two FMAs (exact product inside the FMA, so it stays accurate for any
//               float-valued argument below 2^40) + the fdlibm minimax polynomials on [-pi/4, pi/4] (Sun Microsystems,
//               s_sin.c / k_sin.c / k_cos.c: published constants) evaluated by Estrin's scheme: 4 dependent FMAs
//               instead of 6-7 (Horner).  Returns the quadrant and the reduced argument too.
//   pll_step_fast: the phase detector without a division or an arctangent polynomial.  The detector input is
//               (eI, eQ) = (fl(x*fbI), fl(x*-fbQ)) with (fbI, fbQ) = (fl(cos T), fl(sin T)) from the previous step, so
//               its angle is -T (+pi when x < 0) up to the four float roundings: atan2(eQ, eI) = theta0 + cross/dot
//               with theta0 = -T mod 2pi known to ~1e-16 from the previous step's range reduction, (cross, dot) taken
//               against the previous step's double (cos T, sin T), and |cross/dot| ~ 1e-7 (its cube is below 1e-21).
//               1/dot comes from 1/x (computed off the critical path) and one Newton correction.
// Both are `__host__ __device__` so that tests/ can run the very same code on the CPU against glibc.
#ifndef FMRX_PLLMATH_H
#define FMRX_PLLMATH_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FMRX_HD __host__ __device__ __forceinline__
#define FMRX_HD_COLD inline __host__ __device__ __noinline__
#else
#define FMRX_HD inline
#define FMRX_HD_COLD inline
#endif

namespace fmrx {
namespace pllmath {

FMRX_HD int lo_word(double v) {
#ifdef __CUDA_ARCH__
    return __double2loint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int)(uint32_t)u;
#endif
}

FMRX_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

FMRX_HD int hi_word_(double v) {
#ifdef __CUDA_ARCH__
    return __double2hiint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int)(uint32_t)(u >> 32);
#endif
}
FMRX_HD double make_double_(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}

struct SinCos {
    double sn, cs;  // sin(T), cos(T)
    double r;       // T = n*(pi/2) + r, |r| <~ pi/4
    int q;          // n mod 4
};

// valid for |T| < 2^40 (the caller falls back to libm beyond)
FMRX_HD SinCos sincos_cw(double T) {
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer
    const double P1 = 1.5707963267948966;      // pi/2 rounded to double
    const double P2 = 6.123233995736766e-17;   // pi/2 - P1
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double t = fma_(T, TWO_OVER_PI, MAGIC);
    const double fn = t - MAGIC;
    SinCos o;
    o.q = lo_word(t) & 3;
    double r = fma_(-fn, P1, T);
    r = fma_(-fn, P2, r);
    o.r = r;
    const double z = r * r;
    const double z2 = z * z;
    const double sa = fma_(z, S2, S1), sb = fma_(z, S4, S3), sc = fma_(z, S6, S5);
    const double ca = fma_(z, C2, C1), cb = fma_(z, C4, C3), cc = fma_(z, C6, C5);
    const double z4 = z2 * z2;
    const double rz = r * z;
    const double hz = fma_(z, -0.5, 1.0);
    const double sp = fma_(z4, sc, fma_(z2, sb, sa));
    const double cp = fma_(z4, cc, fma_(z2, cb, ca));
    const double s = fma_(rz, sp, r);   // sin r
    const double c = fma_(z2, cp, hz);  // cos r
    // quadrant: swap for odd n, then the signs -- sin is negative for q in {2, 3}, cos for q in {1, 2}.  The negation is a
    // flip of the sign bit of the high word (one integer op on the dependency chain instead of a DADD and two selects)
    const double sn0 = (o.q & 1) ? c : s, cs0 = (o.q & 1) ? s : c;
    o.sn = make_double_(hi_word_(sn0) ^ (int)(((unsigned)o.q & 2u) << 30), lo_word(sn0));
    o.cs = make_double_(hi_word_(cs0) ^ (int)((((unsigned)o.q + 1u) & 2u) << 30), lo_word(cs0));
    return o;
}

FMRX_HD int hi_word(double v) {
#ifdef __CUDA_ARCH__
    return __double2hiint(v);
#else
    uint64_t u;
    memcpy(&u, &v, 8);
    return (int)(uint32_t)(u >> 32);
#endif
}
FMRX_HD double make_double(int hi, int lo) {
#ifdef __CUDA_ARCH__
    return __hiloint2double(hi, lo);
#else
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v;
    memcpy(&v, &u, 8);
    return v;
#endif
}
FMRX_HD float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
    return __fdividef(1.0f, x);  // MUFU.RCP: ~1e-7 relative, branch-free; its error enters the angle squared
#else
    return 1.0f / x;
#endif
}

// cos(T) alone (the NCO output), same reduction and polynomials
FMRX_HD double cos_cw(double T) {
    const SinCos v = sincos_cw(T);
    return v.cs;
}

// The same routine with its fifteen non-trivial double constants taken from a struct: ptxas treats a 64-bit immediate as free to
// rebuild and re-materialises all of them (two moves each) at the top of every loop iteration, 14 of 176 instructions per
// step; values the kernel has LOADED (from shared memory, once per launch) stay in registers instead.  Same operations in the
// same order as sincos_cw.
struct PllK {
    double two_over_pi, p1, p2, s1, s2, s3, s4, s5, s6, c1, c2, c3, c4, c5, c6;
    unsigned bias_hi;          // 896 << 20, the exponent re-bias of the float -> double re-packing
};
constexpr int kPllKDoubles = 15;
FMRX_HD double pll_k_value(int i) {
    switch (i) {
        case 0: return 6.36619772367581382433e-01;
        case 1: return 1.5707963267948966;
        case 2: return 6.123233995736766e-17;
        case 3: return -1.66666666666666324348e-01;
        case 4: return 8.33333333332248946124e-03;
        case 5: return -1.98412698298579493134e-04;
        case 6: return 2.75573137070700676789e-06;
        case 7: return -2.50507602534068634195e-08;
        case 8: return 1.58969099521155010221e-10;
        case 9: return 4.16666666666666019037e-02;
        case 10: return -1.38888888888741095749e-03;
        case 11: return 2.48015872894767294178e-05;
        case 12: return -2.75573143513906633035e-07;
        case 13: return 2.08757232129817482790e-09;
        default: return -1.13596475577881948265e-11;
    }
}
FMRX_HD PllK pll_k_from(const double *v, unsigned bias_hi) {
    PllK k;
    k.two_over_pi = v[0]; k.p1 = v[1]; k.p2 = v[2];
    k.s1 = v[3]; k.s2 = v[4]; k.s3 = v[5]; k.s4 = v[6]; k.s5 = v[7]; k.s6 = v[8];
    k.c1 = v[9]; k.c2 = v[10]; k.c3 = v[11]; k.c4 = v[12]; k.c5 = v[13]; k.c6 = v[14];
    k.bias_hi = bias_hi;
    return k;
}
FMRX_HD PllK pll_k_literal() {
    double v[kPllKDoubles];
    for (int i = 0; i < kPllKDoubles; ++i) v[i] = pll_k_value(i);
    return pll_k_from(v, 0x38000000u);
}

FMRX_HD SinCos sincos_k(double T, const PllK &K) {
    const double MAGIC = 6755399441055744.0;
    const double t = fma_(T, K.two_over_pi, MAGIC);
    const double fn = t - MAGIC;
    SinCos o;
    o.q = lo_word(t) & 3;
    double r = fma_(-fn, K.p1, T);
    r = fma_(-fn, K.p2, r);
    o.r = r;
    const double z = r * r;
    const double z2 = z * z;
    const double sa = fma_(z, K.s2, K.s1), sb = fma_(z, K.s4, K.s3), sc = fma_(z, K.s6, K.s5);
    const double ca = fma_(z, K.c2, K.c1), cb = fma_(z, K.c4, K.c3), cc = fma_(z, K.c6, K.c5);
    const double z4 = z2 * z2;
    const double rz = r * z;
    const double hz = fma_(z, -0.5, 1.0);
    const double sp = fma_(z4, sc, fma_(z2, sb, sa));
    const double cp = fma_(z4, cc, fma_(z2, cb, ca));
    const double s = fma_(rz, sp, r);
    const double c = fma_(z2, cp, hz);
    const double sn0 = (o.q & 1) ? c : s, cs0 = (o.q & 1) ? s : c;
    o.sn = make_double_(hi_word_(sn0) ^ (int)(((unsigned)o.q & 2u) << 30), lo_word(sn0));
    o.cs = make_double_(hi_word_(cs0) ^ (int)((((unsigned)o.q + 1u) & 2u) << 30), lo_word(cs0));
    return o;
}

// the bare MUFU.RCP (flush-to-zero form): __fdividef wraps it in a range check and two scalings that `ok` makes redundant -- a step
// whose |x| is outside (1e-18, 1e18) is redone by libm whatever this returns
FMRX_HD float rcp_bare(float x) {
#ifdef __CUDA_ARCH__
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

FMRX_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;  // volatile: no contraction with a following add on the host either
    return r;
#endif
}
FMRX_HD float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
FMRX_HD double dmul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
FMRX_HD double dadd_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// one step of fmPLL's loop (src/helper.cpp:32-45) with the reference's operand types: fp32 loop
// state with one rounding per operation, double transcendentals on float-valued arguments, trigArg rounded to fp32.
//
// pll_step_fast is branch-free (one basic block, so that the compiler can interleave the three independent chains of
// a step: phase detector -> loop filter -> oscillator, the NCO output's cosine, and the next sample's reciprocal) and
// reports through `ok` whether its assumptions held; when they did not (first step after loading the state, x zero /
// non-finite / tiny, the detector angle within 1e-5 of the +-pi seam, arguments beyond 2^40) the caller restores the
// four carried floats and redoes the step with pll_step_libm.
// ---------------------------------------------------------------------------------------------------------------
struct PllCarry {  // pll_state_type minus trigOffset / ncoLast (src/helper.h:17-19)
    float integ, phase, fbi, fbq;
};
struct PllCoef {
    float Ki, Kp, scale, adj;
    double w;  // (2*PI) * (double)(freq/Fs)
};
// what the fast detector needs about the trigArg T that produced fbi/fbq: cos T, sin T in double, and
// theta0 = -T mod 2pi (x > 0) or pi - T mod 2pi (x < 0), wrapped into (-pi, pi], as hi + s1 with hi = k*(pi/2).
// Separate scalars: an array indexed by the sign of x would live in local memory.
struct PllFast {
    double cs, sn;
    double th_hi_p, th_s1_p, th_hi_n, th_s1_n;
    bool usable_p, usable_n;  // armed and not within 1e-5 of the seam, for x > 0 / x < 0
    bool neg;                 // V = 1: the sign of x the (single) theta0 in th_hi_p / th_s1_p / usable_p was prepared for
};

constexpr float kTrigLimitF = 1.0e12f;  // < 2^40

FMRX_HD void pll_disarm(PllFast &f) { f.usable_p = f.usable_n = false; }

FMRX_HD void pll_prepare(PllFast &f, const SinCos &v) {
    const double H1 = 1.5707963267948966, L1 = 6.123233995736766e-17;
    f.cs = v.cs;
    f.sn = v.sn;
    const bool rneg = v.r < 0.0;
    const bool near0 = fabs(v.r) < 1e-5;
    // k such that theta0 = k*(pi/2) - r lies in (-pi, pi]: x > 0: q=0 -> 0, 1 -> -1, 2 -> +2 (r >= 0) or -2 (r < 0), 3 -> 1;
    // x < 0 adds pi, i.e. the same table two quadrants on.  Looked up from packed nibbles to stay branch-free.
    const int q = v.q;
    const unsigned kTable = 0x1EF012F0u;  // nibble q (r >= 0), nibble 4+q (r < 0), two's complement
    const int sh = rneg ? 16 : 0;
    const int kp = (int)(((kTable >> (sh + 4 * q)) & 15u) ^ 8u) - 8;
    const int kn = (int)(((kTable >> (sh + 4 * ((q + 2) & 3))) & 15u) ^ 8u) - 8;
    f.th_hi_p = (double)kp * H1; f.th_s1_p = fma_((double)kp, L1, -v.r);
    f.th_hi_n = (double)kn * H1; f.th_s1_n = fma_((double)kn, L1, -v.r);
    f.usable_p = !(q == 2 && near0);
    f.usable_n = !(q == 0 && near0);
}

// ---------------------------------------------------------------------------------------------------------------
// Variants 1 and 2 of the step (round 2, second session): the same arithmetic with fewer conversion and FP64 instructions and
// a shorter dependency chain (488 -> 412 cycles per step for one warp alone; DESIGN 3.4 for what it does and does not buy on
// the partition).  A float -> double conversion of a NORMAL finite float is a re-packing of its bits (sign, exponent + 896,
// mantissa << 29) -- exact, three integer instructions and a shift instead of an F2F.F64.F32 on the 16-lane conversion pipe;
// `ok` already sends zero / non-finite values to the libm path and now also anything outside [FLT_MIN, 1e30).  theta0 is
// prepared for the sign of the NEXT sample only (the caller knows it: the input is in registers two groups ahead), with
// k * (pi/2) and k * (pi/2's tail) read from a 16-entry table (k is -2..2) instead of two int -> double conversions, two
// multiplies and one of the two FMAs; a wrong prediction is caught in `ok` (f.neg) and costs one libm group.
// ---------------------------------------------------------------------------------------------------------------
FMRX_HD unsigned float_bits(float v) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(v);
#else
    unsigned u;
    memcpy(&u, &v, 4);
    return u;
#endif
}
// |v| in [FLT_MIN, 1e30): false for zero, subnormal, huge, infinite and NaN
FMRX_HD bool is_plain(float v) { return fabsf(v) >= 1.17549435e-38f && fabsf(v) < 1.0e30f; }
// (double)v for a normal finite float, exact
FMRX_HD double widen(float v, const PllK &K) {
    // arithmetic shift: ssss eeeeeeee mmm...; drop the three sign copies below the top one, re-bias the exponent
    const unsigned b = float_bits(v);
    const unsigned h = (unsigned)((int)b >> 3);
    return make_double((int)((h & 0x8FFFFFFFu) + K.bias_hi), (int)(b << 29));
}
// the same for a value known to be positive
FMRX_HD double widen_pos(float v, const PllK &K) {
    const unsigned t = float_bits(v);
    return make_double((int)((t >> 3) + K.bias_hi), (int)(t << 29));
}

// theta0 = k*(pi/2) - r in (-pi, pi] as (k*H1, k*L1), H1 + L1 = pi/2: looked up by (quadrant of T, sign of r, sign of the
// next sample) -- sixteen pairs of doubles; the kernel keeps the table in shared memory (one LDS.128 per step on a pipe the
// step does not otherwise use), the host builds the entry on the fly.
//   qq = (q + 2*[x < 0]) & 3:  0 -> k = 0, 1 -> -1, 2 -> +2 (r >= 0) or -2 (r < 0), 3 -> +1
struct PllTheta {
    double kh, kl;
};
FMRX_HD int pll_theta_index(int q, bool rneg, bool neg_next) { return q | (rneg ? 4 : 0) | (neg_next ? 8 : 0); }
FMRX_HD PllTheta pll_theta_entry(int idx) {
    const double H1 = 1.5707963267948966, L1 = 6.123233995736766e-17;
    const int qq = ((idx & 3) + ((idx & 8) ? 2 : 0)) & 3;
    const bool rneg = (idx & 4) != 0;
    const double k = qq == 0 ? 0.0 : qq == 1 ? -1.0 : qq == 3 ? 1.0 : (rneg ? -2.0 : 2.0);
    PllTheta t;
    t.kh = k * H1;  // exact: k is 0, +-1, +-2
    t.kl = k * L1;
    return t;
}

FMRX_HD void pll_prepare1(PllFast &f, const SinCos &v, bool neg_next, const PllTheta *table) {
    f.cs = v.cs;
    f.sn = v.sn;
    const int rh = hi_word(v.r);
    const bool rneg = rh < 0;                                      // sign bit (a -0 is inside the seam band anyway)
    const bool near0 = (unsigned)(rh & 0x7fffffff) < 0x3EE4F8B5u;  // |r| < ~1e-5
    const int idx = pll_theta_index(v.q, rneg, neg_next);
#ifdef __CUDA_ARCH__
    const PllTheta t = table[idx];
#else
    const PllTheta t = table ? table[idx] : pll_theta_entry(idx);
#endif
    f.th_hi_p = t.kh;
    f.th_s1_p = t.kl - v.r;  // = fma(k, L1, -r): k * L1 is exact
    const int q2 = neg_next ? 0 : 2;  // the quadrant whose theta0 sits on the +-pi seam
    f.usable_p = !((v.q == q2) & near0);
    f.neg = neg_next;
}

// re-arm the fast detector from the float trigArg of the step just taken
FMRX_HD void pll_rearm(PllFast &f, float trig) {
    if (fabsf(trig) < kTrigLimitF) pll_prepare(f, sincos_cw((double)trig));
    else pll_disarm(f);
}

FMRX_HD void pll_rearm1(PllFast &f, float trig, bool neg_next, const PllTheta *table, const PllK &K) {
    if (is_plain(trig) && fabsf(trig) < kTrigLimitF) pll_prepare1(f, sincos_k((double)trig, K), neg_next, table);
    else pll_disarm(f);
}

// x: input sample; cnt = (trigOffset + k) + 1 as the reference forms it in fp32; returns nco[k+1]
FMRX_HD float pll_step_fast(PllCarry &c, PllFast &f, const PllCoef &p, float x, float cnt, bool &ok) {
    const bool neg = x < 0.0f;
    const float ax = fabsf(x);
    const float eI = mul_rn(x, c.fbi);
    const float eQ = mul_rn(x, -c.fbq);
    ok = (neg ? f.usable_n : f.usable_p) && ax > 1e-18f && ax < 1e18f && eI != 0.0f && eQ != 0.0f;
    const double rx = (double)rcp_approx(x);
    const double dI = (double)eI, dQ = (double)eQ;
    const double dot = fma_(dQ, -f.sn, dI * f.cs);
    const double cross = fma_(dQ, f.cs, dI * f.sn);
    const double e = fma_(-dot, rx, 2.0);
    const double raw = (neg ? f.th_hi_n : f.th_hi_p) + fma_(cross * rx, e, neg ? f.th_s1_n : f.th_s1_p);
    const double ang = make_double((hi_word(raw) & 0x7fffffff) | (hi_word(dQ) & 0x80000000), lo_word(raw));  // copysign(|raw|, eQ)
    const float eD = (float)ang;
    c.integ = add_rn(c.integ, mul_rn(p.Ki, eD));
    c.phase = add_rn(c.phase, add_rn(mul_rn(p.Kp, eD), c.integ));
    const float trig = (float)dadd_rn(dmul_rn(p.w, (double)cnt), (double)c.phase);  // src/helper.cpp:41: double expression, no FMA
    const float targ = add_rn(mul_rn(trig, p.scale), p.adj);
    ok = ok && fabsf(trig) < kTrigLimitF && fabsf(targ) < kTrigLimitF;
    const SinCos v = sincos_cw((double)trig);
    pll_prepare(f, v);
    c.fbi = (float)v.cs;
    c.fbq = (float)v.sn;
    return (float)cos_cw((double)targ);
}

// variant 1 (see above): neg_next = sign of the sample the NEXT step will see
// F2F_SIDE: the two widenings off the detector -> loop filter -> oscillator chain (1/x and the NCO argument) as conversion
// instructions again: 6 instructions fewer per step for 16 more cycles of the conversion pipe (kernel variant 2)
template <bool F2F_SIDE = false>
FMRX_HD float pll_step_fast1(PllCarry &c, PllFast &f, const PllCoef &p, const PllK &K, float x, float cnt, bool neg_next, const PllTheta *table, bool &ok) {
    const bool neg = x < 0.0f;
    const float ax = fabsf(x);
    const float eI = mul_rn(x, c.fbi);
    const float eQ = mul_rn(x, -c.fbq);
    // `&`, not `&&`: one basic block.  x in (1e-18, 1e18) and |fbi|, |fbq| <= 1 bound eI, eQ from above; cnt = (offset + k) + 1
    // with an offset the caller has checked is plain once per block
    bool good = f.usable_p & (neg == f.neg) & (ax > 1e-18f) & (ax < 1e18f) & (fabsf(eI) >= 1.17549435e-38f) & (fabsf(eQ) >= 1.17549435e-38f);
    const double rx = F2F_SIDE ? (double)rcp_bare(x) : widen(rcp_bare(x), K);
    const double dI = widen(eI, K), dQ = widen(eQ, K);
    const double dot = fma_(dQ, -f.sn, dI * f.cs);
    const double cross = fma_(dQ, f.cs, dI * f.sn);
    const double e = fma_(-dot, rx, 2.0);
    const double raw = f.th_hi_p + fma_(cross * rx, e, f.th_s1_p);
    const double ang = make_double((hi_word(raw) & 0x7fffffff) | (int)(float_bits(eQ) & 0x80000000u), lo_word(raw));  // copysign(|raw|, eQ)
    const float eD = (float)ang;
    c.integ = add_rn(c.integ, mul_rn(p.Ki, eD));
    c.phase = add_rn(c.phase, add_rn(mul_rn(p.Kp, eD), c.integ));
    const float trig = (float)dadd_rn(dmul_rn(p.w, widen_pos(cnt, K)), widen(c.phase, K));  // src/helper.cpp:41: double expression, no FMA
    const float targ = add_rn(mul_rn(trig, p.scale), p.adj);
    good = good & is_plain(c.phase) & (fabsf(trig) >= 1.17549435e-38f) & (fabsf(trig) < kTrigLimitF) & (fabsf(targ) >= 1.17549435e-38f) & (fabsf(targ) < kTrigLimitF);
    ok = good;
    const SinCos v = sincos_k(widen(trig, K), K);
    pll_prepare1(f, v, neg_next, table);
    c.fbi = (float)v.cs;
    c.fbq = (float)v.sn;
    return (float)sincos_k(F2F_SIDE ? (double)targ : widen(targ, K), K).cs;
}

// the same step through libm, everything by value (a cold, out-of-line call on the device must not pin the caller's
// state in local memory)
struct PllLibmOut {
    PllCarry c;
    float nco, trig;
};
FMRX_HD_COLD PllLibmOut pll_step_libm(PllCarry c, PllCoef p, float x, float cnt) {
    const float eI = mul_rn(x, c.fbi);
    const float eQ = mul_rn(x, -c.fbq);
    const float eD = (float)atan2((double)eQ, (double)eI);
    c.integ = add_rn(c.integ, mul_rn(p.Ki, eD));
    c.phase = add_rn(c.phase, add_rn(mul_rn(p.Kp, eD), c.integ));
    const float trig = (float)dadd_rn(dmul_rn(p.w, (double)cnt), (double)c.phase);
    c.fbi = (float)cos((double)trig);
    c.fbq = (float)sin((double)trig);
    PllLibmOut o;
    o.c = c;
    o.trig = trig;
    o.nco = (float)cos((double)add_rn(mul_rn(trig, p.scale), p.adj));
    return o;
}

// host-side convenience (tests): one step with the per-step fallback
FMRX_HD float pll_step(PllCarry &c, PllFast &f, const PllCoef &p, float x, float cnt) {
    const PllCarry saved = c;
    bool ok;
    float out = pll_step_fast(c, f, p, x, cnt, ok);
    if (!ok) {
        const PllLibmOut o = pll_step_libm(saved, p, x, cnt);
        c = o.c;
        out = o.nco;
        pll_rearm(f, o.trig);
    }
    return out;
}

FMRX_HD float pll_step1(PllCarry &c, PllFast &f, const PllCoef &p, float x, float cnt, bool neg_next, const PllTheta *table) {
    const PllK K = pll_k_literal();
    const PllCarry saved = c;
    bool ok;
    float out = pll_step_fast1(c, f, p, K, x, cnt, neg_next, table, ok);
    if (!ok || !is_plain(cnt)) {
        const PllLibmOut o = pll_step_libm(saved, p, x, cnt);
        c = o.c;
        out = o.nco;
        pll_rearm1(f, o.trig, neg_next, table, K);
    }
    return out;
}

}  // namespace pllmath
}  // namespace fmrx
#endif
