// CPU check of the PLL's double-precision kernels (the product header fmrx_pllmath.h compiled for the host) against
// glibc, which is what the reference links (fmPLL, /root/reference/src/helper.cpp:13-57 calls atan2/cos/sin on doubles).
//   pllmath_check sincos <n>        max error of sincos_cw vs glibc sin/cos in double ulps over float-valued arguments
//   pllmath_check loop <blocks> [variant]   runs the fast loop (variant 0: conversion instructions; 1: integer-built
//                                   conversions, theta0 for the next sample's sign) and a libm-only loop side by side and counts float mismatches
//   pllmath_check widen <n>         the integer-built float -> double conversion against the cast
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "fmrx_pllmath.h"

using namespace fmrx::pllmath;

static double ulps(double got, double want) {
    if (got == want) return 0.0;
    int e;
    frexp(want, &e);
    return fabs(got - want) / ldexp(1.0, e - 53);
}

// libm-only restatement of the same recurrence (the shape of oracle/fmrx_oracle.c's PLL)
struct RefLoop { float integ, phase, fbi, fbq, Ki, Kp, scale, adj; double w; };
static float ref_step(RefLoop &c, float x, float cnt) {
    const float eI = mul_rn(x, c.fbi), eQ = mul_rn(x, -c.fbq);
    const float eD = (float)atan2((double)eQ, (double)eI);
    c.integ = add_rn(c.integ, mul_rn(c.Ki, eD));
    c.phase = add_rn(c.phase, add_rn(mul_rn(c.Kp, eD), c.integ));
    const float trig = (float)(c.w * (double)cnt + (double)c.phase);
    c.fbi = (float)cos((double)trig);
    c.fbq = (float)sin((double)trig);
    return (float)cos((double)add_rn(mul_rn(trig, c.scale), c.adj));
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::mt19937_64 rng(12345);
    if (!strcmp(argv[1], "sincos")) {
        const long n = atol(argv[2]);
        double worst_s = 0, worst_c = 0, worst_abs = 0;
        std::uniform_real_distribution<double> mag(-3.0, 9.0);
        long flips = 0;
        for (long i = 0; i < n; ++i) {
            float T = (float)(pow(10.0, mag(rng)) * ((rng() & 1) ? 1 : -1));
            const SinCos v = sincos_cw((double)T);
            const double s = sin((double)T), c = cos((double)T);
            worst_s = fmax(worst_s, ulps(v.sn, s));
            worst_c = fmax(worst_c, ulps(v.cs, c));
            worst_abs = fmax(worst_abs, fmax(fabs(v.sn - s), fabs(v.cs - c)));
            flips += ((float)v.sn != (float)s) + ((float)v.cs != (float)c);
        }
        printf("{\"n\": %ld, \"max_ulp_sin\": %.3f, \"max_ulp_cos\": %.3f, \"max_abs\": %.3e, \"float_flips\": %ld}\n", n, worst_s, worst_c, worst_abs, flips);
        return 0;
    }
    if (!strcmp(argv[1], "widen")) {
        const long n = atol(argv[2]);
        long bad = 0;
        for (long i = 0; i < n; ++i) {
            unsigned u = (unsigned)rng();
            float v;
            memcpy(&v, &u, 4);
            if (!is_plain(v)) continue;
            const PllK K = pll_k_literal();
            const double w = widen(v, K), wp = widen_pos(fabsf(v), K);
            bad += (w != (double)v) || (wp != (double)fabsf(v)) || (std::signbit(w) != std::signbit(v));
        }
        const float edge[] = {1.17549435e-38f, -1.17549435e-38f, 9.99999e29f, 1.0f, -1.0f, 16777216.0f, 3.0e-20f};
        for (float v : edge) bad += widen(v, pll_k_literal()) != (double)v;
        printf("{\"n\": %ld, \"bad\": %ld}\n", n, bad);
        return 0;
    }
    if (!strcmp(argv[1], "loop")) {
        const int blocks = atoi(argv[2]), N = 15360, variant = argc > 3 ? atoi(argv[3]) : 0;
        std::vector<float> xs(N + 1);
        PllTheta table[16];  // variant 2: the lookup table as the kernel stages it
        for (int i = 0; i < 16; ++i) table[i] = pll_theta_entry(i);
        long total = 0, mism = 0, fast_steps = 0;
        double max_nco_diff = 0;
        for (int cfg = 0; cfg < 2; ++cfg) {
            const float freq = cfg ? 114000.0f : 19e3f, Fs = 240e3f, scale = cfg ? 0.5f : 2.0f, bw = cfg ? 0.001f : 0.01f;
            const float adj = cfg ? (float)((double)(float)(3.14159265358979323846 / 3.3 - 3.14159265358979323846 / 1.5) - 3.14159265358979323846 / 1.4) : 0.0f;
            for (int trial = 0; trial < 4; ++trial) {
                PllCarry c{0, 0, 1, 0};
                PllFast f{};
                pll_disarm(f);
                PllCoef p{(bw * bw) * 3.555f, bw * 2.666f, scale, adj, (2 * 3.14159265358979323846) * (double)(freq / Fs)};
                RefLoop r{0, 0, 1, 0, p.Ki, p.Kp, scale, adj, p.w};
                float off = 0.0f;
                std::normal_distribution<double> noise(0.0, trial == 3 ? 0.5 : 0.01);
                const double amp = trial == 1 ? 1e-3 : trial == 2 ? 30.0 : 0.05, ph0 = 0.3 + trial, df = trial == 2 ? 3.0 : 0.0;
                for (int b = 0; b < blocks; ++b) {
                    pll_disarm(f);  // a launch boundary: only the float state is carried
                    for (int k = 0; k < N; ++k) {
                        const double t = ((double)b * N + k) / Fs;
                        float x = (float)(amp * cos(2 * 3.14159265358979323846 * (freq + df) * t + ph0) + amp * noise(rng));
                        if (trial == 3 && (k % 977) == 0) x = 0.0f;  // exact zeros exercise the libm path
                        xs[k] = x;
                    }
                    xs[N] = 1.0f;  // the kernel does not look across a block boundary either: it predicts "positive"
                    for (int k = 0; k < N; ++k) {
                        const float x = xs[k];
                        const bool neg_next = xs[k + 1] < 0.0f;
                        const float cnt = add_rn(add_rn(off, (float)k), 1.0f);
                        {
                            bool ok; PllCarry pc = c; PllFast pf = f;
                            if (variant) pll_step_fast1(pc, pf, p, pll_k_literal(), x, cnt, neg_next, variant == 2 ? table : nullptr, ok); else pll_step_fast(pc, pf, p, x, cnt, ok);
                            fast_steps += ok;
                        }
                        const float a = variant ? pll_step1(c, f, p, x, cnt, neg_next, variant == 2 ? table : nullptr) : pll_step(c, f, p, x, cnt), g = ref_step(r, x, cnt);
                        ++total;
                        const bool same = a == g && c.integ == r.integ && c.phase == r.phase && c.fbi == r.fbi && c.fbq == r.fbq;
                        if (!same) {
                            ++mism;
                            max_nco_diff = fmax(max_nco_diff, fabs((double)a - (double)g));
                            // resynchronise so that one flip is counted once
                            c.integ = r.integ; c.phase = r.phase; c.fbi = r.fbi; c.fbq = r.fbq; pll_disarm(f);
                        }
                    }
                    off = add_rn(off, (float)N);
                }
            }
        }
        printf("{\"steps\": %ld, \"fast_steps\": %ld, \"mismatches\": %ld, \"max_nco_diff\": %.3e}\n", total, fast_steps, mism, max_nco_diff);
        return 0;
    }
    return 2;
}
