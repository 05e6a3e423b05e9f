// TEST INFRASTRUCTURE ONLY — never linked into, imported by, or called from the product.
//
// extern "C" shim over the UNMODIFIED reference objects (compiled by oracle/Makefile from the
// sources where they lie under /root/reference/src; outputs only into oracle/_ref/).
// Each entry wraps caller-owned C arrays into the std::vector / float*& arguments the reference
// free functions take (src/filter.h:17-38, src/helper.h:21-25, src/rf_module.h:20,
// src/iofunc.h:28) and copies the results back.  No arithmetic happens in this file.
//
// The reference accumulates into its output vectors (`y[n] += ...`); every wrapper
// therefore takes an optional `y_init` so a test can reproduce either the fresh-vector or the
// accumulate-across-blocks behaviour.
#include <cstdint>
#include <cstring>
#include <condition_variable>
#include <iostream>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include "filter.h"
#include "helper.h"
#include "iofunc.h"
#include "rf_module.h"

// thread bodies of src/fm_radio.cpp (non-static free functions; declared here, defined in fm_radio.o)
void frame_thread(int &mode, std::queue<std::vector<float>> &frame_queue, std::queue<void *> &rds_queue,
                  std::mutex &frame_mutex, std::condition_variable &cvarframe);

namespace {
std::vector<float> vec(const float *p, size_t n) { return p ? std::vector<float>(p, p + n) : std::vector<float>(n, 0.0f); }
void out(const std::vector<float> &v, float *p, size_t cap) {
    if (p) std::memcpy(p, v.data(), sizeof(float) * (v.size() < cap ? v.size() : cap));
}
void load_pll(pll_state_type &s, const float *st) {
    s.integrator = st[0]; s.phaseEst = st[1]; s.feedbackI = st[2]; s.feedbackQ = st[3]; s.trigOffset = st[4]; s.ncoLast = st[5];
}
void store_pll(const pll_state_type &s, float *st) {
    st[0] = s.integrator; st[1] = s.phaseEst; st[2] = s.feedbackI; st[3] = s.feedbackQ; st[4] = s.trigOffset; st[5] = s.ncoLast;
}
}  // namespace

extern "C" {

// src/filter.cpp:19-38 / 41-60 / 63-93
void ref_lpf(float Fs, float Fc, unsigned short ntaps, float *h) { std::vector<float> v; impulseResponseLPF(Fs, Fc, ntaps, v); out(v, h, ntaps); }
void ref_bpf(float Fb, float Fe, float Fs, int ntaps, float *h) { std::vector<float> v; impulseResponseBPF(Fb, Fe, Fs, ntaps, v); out(v, h, ntaps); }
void ref_rrc(float Fs, int ntaps, float *h) { std::vector<float> v; impulseResponseRRC(Fs, ntaps, v); out(v, h, ntaps); }

// src/iofunc.cpp:61-69 — feeds std::cin from memory; a short `navail` reproduces the short-read padding
void ref_unpack(const uint8_t *raw, unsigned navail, unsigned nsamples, float *dst) {
    std::istringstream iss(std::string(reinterpret_cast<const char *>(raw), navail));
    std::streambuf *old = std::cin.rdbuf(iss.rdbuf());
    std::cin.clear();
    std::vector<float> block(nsamples, 0.0f);
    unsigned id = 0;
    readStdInBlock(nsamples, id, block);
    std::cin.rdbuf(old);
    std::cin.clear();
    out(block, dst, nsamples);
}

// src/filter.cpp:126-154
int ref_conv_decim(const float *y_init, float *y, const float *x, int n, const float *h, int nt, float *zi, int decim) {
    std::vector<float> yv = vec(y_init, y_init ? n / decim : 0), xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nt - 1);
    convolveWithDecim(yv, xv, hv, ziv, decim);
    out(yv, y, yv.size()); out(ziv, zi, nt - 1);
    return (int)yv.size();
}
// src/filter.cpp:157-185
int ref_conv_decim_ptr(const float *y_init, float *y, const float *x, int n, const float *h, int nt, float *zi, int nzi, int decim) {
    std::vector<float> yv = vec(y_init, y_init ? n / decim : 0), xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nzi);
    float *xp = xv.data();
    convolveWithDecimPointer(yv, xp, (unsigned)n, hv, ziv, decim);
    out(yv, y, yv.size()); out(ziv, zi, nzi);
    return (int)yv.size();
}
// src/filter.cpp:187-219
int ref_conv_decim_iq(float *yi, float *yq, const float *xi, const float *xq, int n, const float *h, int nt, float *zii, float *ziq, int decim) {
    std::vector<float> a, b, xa(xi, xi + n), xb(xq, xq + n), hv(h, h + nt), za(zii, zii + nt - 1), zb(ziq, ziq + nt - 1);
    convolveWithDecimIQ(a, xa, hv, za, b, xb, zb, decim);
    out(a, yi, a.size()); out(b, yq, b.size()); out(za, zii, nt - 1); out(zb, ziq, nt - 1);
    return (int)a.size();
}
// src/filter.cpp:222-259 ; `ny_keep` limits how much of y is copied back (mode-1 stereo computes 73728, reads 2949)
int ref_conv_mode1(float *y, int ny_keep, const float *x, int n, const float *h, int nt, float *zi, int nzi, int decim, int up) {
    std::vector<float> yv, xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nzi);
    convolveWithDecimMode1(yv, xv, hv, ziv, decim, up);
    out(yv, y, ny_keep); out(ziv, zi, nzi);
    return (int)yv.size();
}
// src/filter.cpp:261-298
int ref_conv_mode1_ptr(float *y, const float *x, int n, const float *h, int nt, float *zi, int nzi, int decim, int up) {
    std::vector<float> yv, xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nzi);
    float *xp = xv.data();
    convolveWithDecimMode1Pointer(yv, xp, (unsigned)n, hv, ziv, decim, up);
    out(yv, y, yv.size()); out(ziv, zi, nzi);
    return (int)yv.size();
}
// src/filter.cpp:301-339
int ref_conv_mode1_rds(float *y, const float *x, int n, const float *h, int nt, float *zi, int nzi, int decim, int up) {
    std::vector<float> yv, xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nzi);
    convolveWithDecimMode1RDS(yv, xv, hv, ziv, decim, up);
    out(yv, y, yv.size()); out(ziv, zi, nzi);
    return (int)yv.size();
}
// src/filter.cpp:373-401.  The reference sizes everything from x (the untrimmed 15361-long NCO vector) and reads
// x1[x.size()-1] one past the end of the 15360-long x1 (Q8).  Here x1 is given n slots with the last one = `x1_pad`
// so the read is defined; the element it lands in (y[n-1]) is never consumed downstream.
int ref_conv_mixer(float *y, const float *x, const float *x1, int n, int n1, float x1_pad, const float *h, int nt, float *zi, int decim) {
    std::vector<float> yv, xv(x, x + n), x1v(n, x1_pad), hv(h, h + nt), ziv(zi, zi + nt - 1);
    std::memcpy(x1v.data(), x1, sizeof(float) * (n1 < n ? n1 : n));
    convolveWithDecimAndMixer(yv, xv, x1v, hv, ziv, decim);
    out(yv, y, yv.size()); out(ziv, zi, nt - 1);
    return (int)yv.size();
}
// src/rf_module.cpp:13-34
void ref_demod(const float *I, const float *Q, int n, float *dst) {
    std::vector<float> iv(I, I + n), qv(Q, Q + n), prev(2, 0.0f);
    float *p = dst;
    fmDemodArctan(iv, qv, prev, p);
}
// src/helper.cpp:13-57 ; st = {integrator, phaseEst, feedbackI, feedbackQ, trigOffset, ncoLast}
int ref_fmpll(float *nco, const float *x, int n, float freq, float Fs, float scale, float phase_adj, float bw, float *st) {
    std::vector<float> o, xv(x, x + n);
    pll_state_type s; load_pll(s, st);
    fmPLL(o, xv, freq, Fs, scale, phase_adj, bw, s);
    store_pll(s, st); out(o, nco, o.size());
    return (int)o.size();
}
// src/helper.cpp:108-173 ; nco gets n+1 elements (not trimmed by the reference)
int ref_pll_combine(float *y, float *nco, const float *x, int n, const float *h, int nt, float *zi, int decim,
                    float freq, float Fs, float scale, float phase_adj, float bw, float *st) {
    std::vector<float> yv, o, xv(x, x + n), hv(h, h + nt), ziv(zi, zi + nt - 1);
    pll_state_type s; load_pll(s, st);
    pllCombine(yv, xv, hv, ziv, decim, o, freq, Fs, scale, phase_adj, bw, s);
    store_pll(s, st); out(yv, y, yv.size()); out(o, nco, o.size()); out(ziv, zi, nt - 1);
    return (int)o.size();
}

// src/fm_radio.cpp:444-729 — runs the real frame_thread body over `nblk` RRC blocks that are queued up front.
// std::cin is put into the EOF state so the thread's exit test (`:723`) fires once the queue drains; std::cerr is
// captured into `log` (NUL-terminated, truncated to cap).  Returns the untruncated length.
int ref_frame_thread(const float *rrc, int nblk, int blk_len, char *log, int cap) {
    std::queue<std::vector<float>> fq;
    std::queue<void *> rq;
    std::mutex m;
    std::condition_variable cv;
    for (int b = 0; b < nblk; ++b) fq.push(std::vector<float>(rrc + (size_t)b * blk_len, rrc + (size_t)(b + 1) * blk_len));
    std::ostringstream cap_err;
    std::streambuf *old_err = std::cerr.rdbuf(cap_err.rdbuf());
    std::ios_base::iostate old_in = std::cin.rdstate();
    std::cin.setstate(std::ios_base::eofbit);
    int mode = 0;
    frame_thread(mode, fq, rq, m, cv);
    std::cin.clear(old_in);
    std::cerr.rdbuf(old_err);
    std::string s = cap_err.str();
    if (log && cap > 0) {
        size_t c = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
        std::memcpy(log, s.data(), c);
        log[c] = 0;
    }
    return (int)s.size();
}

}  // extern "C"
