// Host-side FIR design.  The taps have to be bit-identical to the reference's, so each expression keeps the operand
// types of the reference expression it restates:
//   impulseResponseLPF  /root/reference/src/filter.cpp:19-38
//   impulseResponseBPF  /root/reference/src/filter.cpp:41-60
//   impulseResponseRRC  /root/reference/src/filter.cpp:63-93
// (fp32 normalised frequencies, double libm, result rounded to fp32; window sin^2(i*pi/N); the low-pass centres its
// sinc on the INTEGER N/2 while testing for the centre at (N-1)/2, which makes tap N/2 a NaN for even N — the mode-1
// 3624-tap filter relies on that)
#include <cmath>

#include "fmrx.h"
#include "fmrx_internal.h"

namespace {
constexpr double kPi = 3.14159265358979323846;  // src/dy4.h:13

inline double hann_sq(int i, int n) {
    const double s = std::sin((static_cast<double>(i) * kPi) / static_cast<double>(n));
    return s * s;
}
}  // namespace

extern "C" int fmrx_design_lpf(float Fs, float Fc, unsigned short ntaps, float *h) {
    if (!h || ntaps == 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_design_lpf: bad argument");
    const int n = ntaps;
    const float cutoff = static_cast<float>(static_cast<double>(Fc) / (static_cast<double>(Fs) / 2.0));
    const int centre = (n - 1) / 2, sinc_origin = n / 2;
    for (int i = 0; i < n; ++i) {
        float tap = cutoff;
        if (i != centre) {
            const double a = kPi * static_cast<double>(cutoff) * static_cast<double>(i - sinc_origin);
            tap = cutoff * static_cast<float>(std::sin(a) / a);
        }
        h[i] = static_cast<float>(static_cast<double>(tap) * hann_sq(i, n));
    }
    return FMRX_OK;
}

extern "C" int fmrx_design_bpf(float Fb, float Fe, float Fs, int ntaps, float *h) {
    if (!h || ntaps <= 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_design_bpf: bad argument");
    const float nyq = Fs / 2.0f;
    const float width = (Fe - Fb) / nyq;
    const float centre_f = ((Fe + Fb) / 2.0f) / nyq;
    const float mid = static_cast<float>((ntaps - 1) / 2);
    for (int i = 0; i < ntaps; ++i) {
        const float fi = static_cast<float>(i);
        float tap = width;
        if (fi != mid) {
            const double a = kPi * static_cast<double>(width / 2.0f) * static_cast<double>(fi - mid);
            tap = static_cast<float>(static_cast<double>(width) * std::sin(a) / a);
        }
        tap = static_cast<float>(static_cast<double>(tap) * std::cos(static_cast<double>(i) * kPi * static_cast<double>(centre_f)));
        h[i] = static_cast<float>(static_cast<double>(tap) * hann_sq(i, ntaps));
    }
    return FMRX_OK;
}

extern "C" int fmrx_design_rrc(float Fs, int ntaps, float *h) {
    if (!h || ntaps <= 0) return fmrx::fail(FMRX_ERR_ARG, "fmrx_design_rrc: bad argument");
    const float Tf = static_cast<float>(1.0 / 2375.0);
    const float bf = 0.90f;
    const double T = Tf, beta = bf;
    const double t_sing = T / (4.0 * beta);
    for (int k = 0; k < ntaps; ++k) {
        const float tf = static_cast<float>(static_cast<double>(k) - static_cast<double>(ntaps) / 2.0) / Fs;
        const double t = tf;
        double v;
        if (t == 0.0) {
            v = 1.0 + beta * ((4 / kPi) - 1);
        } else if (t == -t_sing || t == t_sing) {
            const double a = kPi / (4.0 * beta);
            v = (beta / std::sqrt(2.0)) * (((1.0 + 2.0 / kPi) * std::sin(a)) + ((1.0 - 2.0 / kPi) * std::cos(a)));
        } else {
            const double q = 4.0 * beta * t / T;
            const double num = std::sin(kPi * t * (1.0 - beta) / T) + 4.0 * beta * static_cast<double>(tf / Tf) * std::cos(kPi * t * (1.0 + beta) / T);
            v = num / (kPi * t * (1.0 - q * q) / T);
        }
        h[k] = static_cast<float>(v);
    }
    return FMRX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// `quality` profile: what the reference's report proposes and never built.  None of this is in the
// reference's signal path, so there is nothing to be bit-identical with; the arithmetic is double on the host.
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int fmrx_fir_response(const float *h, int ntaps, float Fs, float f, double *mag, double *phase) {
    if (!h || ntaps <= 0 || !(Fs > 0.0f)) return fmrx::fail(FMRX_ERR_ARG, "fmrx_fir_response: bad argument");
    double re = 0.0, im = 0.0;
    const double w = 2.0 * kPi * static_cast<double>(f) / static_cast<double>(Fs);
    for (int k = 0; k < ntaps; ++k) {
        re += static_cast<double>(h[k]) * std::cos(w * k);
        im -= static_cast<double>(h[k]) * std::sin(w * k);
    }
    if (mag) *mag = std::hypot(re, im);
    if (phase) *phase = std::atan2(im, re);
    return FMRX_OK;
}

// impulseResponseBPF scaled to unit gain at the centre of its pass band (the reference's has 0.308 at 19 kHz, src/filter.cpp:41-60:
// the cosine modulation halves a low-pass prototype whose own gain the sin^2 window already lowered)
extern "C" int fmrx_design_bpf_unity(float Fb, float Fe, float Fs, int ntaps, float *h) {
    if (int e = fmrx_design_bpf(Fb, Fe, Fs, ntaps, h)) return e;
    double mag = 0.0;
    fmrx_fir_response(h, ntaps, Fs, (Fb + Fe) / 2.0f, &mag, nullptr);
    if (!(mag > 0.0)) return fmrx::fail(FMRX_ERR_ARG, "fmrx_design_bpf_unity: the filter has no gain at its band centre");
    for (int k = 0; k < ntaps; ++k) h[k] = static_cast<float>(static_cast<double>(h[k]) / mag);
    return FMRX_OK;
}

// The phase the 114 kHz NCO has to be turned by so that the regenerated 57 kHz carrier lines up with the RDS band it is mixed with
// -- what the reference tunes by hand (phaseAdjust = pi/3.3 - pi/1.5, then "- PI/1.4", src/fm_radio.cpp:342,400).  The RDS band
// r = m cos(w n + p) is squared and band-passed (pllCombine's filter, response H at 2w): the loop locks to 2 w n + 2 p + arg H(2w),
// the NCO halves that (ncoScale 0.5), and its output for sample n is the one computed for it (the one-sample delay of ncoOut and the
// "+1" of trigArg cancel): cos(w n + p + arg H(2w) / 2 + adj).  Mixing with r is coherent for adj = -arg H(2w) / 2 (mod pi; the
// polarity is irrelevant behind a differential decoder).
extern "C" int fmrx_rds_auto_phase(const float *h_sq, int ntaps, float Fs, float f2, float *phase_adj) {
    if (!phase_adj) return fmrx::fail(FMRX_ERR_ARG, "fmrx_rds_auto_phase: null pointer");
    double ph = 0.0;
    if (int e = fmrx_fir_response(h_sq, ntaps, Fs, f2, nullptr, &ph)) return e;
    *phase_adj = static_cast<float>(-0.5 * ph);
    return FMRX_OK;
}

// 1 / (1 + s tau) through the bilinear transform at Fs: y[n] = b (x[n] + x[n-1]) - a1 y[n-1]
extern "C" int fmrx_deemphasis_coeffs(float tau_us, float Fs, double *b, double *a1) {
    if (!b || !a1 || !(tau_us > 0.0f) || !(Fs > 0.0f)) return fmrx::fail(FMRX_ERR_ARG, "fmrx_deemphasis_coeffs: bad argument");
    const double k = 2.0 * static_cast<double>(Fs) * static_cast<double>(tau_us) * 1e-6;
    *b = 1.0 / (1.0 + k);
    *a1 = (1.0 - k) / (1.0 + k);
    return FMRX_OK;
}
