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
//                           counter / resync, 27-bit carry.  Integer results are bit-exact by construction: the only
//                           floating-point operations are comparisons of RRC samples.
#include <cuda_runtime.h>

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
__global__ void resample_kernel(const ResDev a) {
    const int s = blockIdx.z, b = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= a.ny) return;
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

__global__ void resample_state_kernel(const ResDev a) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nzi) return;
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

enum { W_BLOCK = 0, W_OFFSET, W_START, W_LONELY, W_FRONT, W_PREBIT, W_NBITS, W_PRINTPOS, W_LASTPOS1, W_BAD, W_CARRY = 10, W_BITS = 40 };

__device__ __forceinline__ bool same_sign(float a, float b) { return (a > 0 && b > 0) || (a < 0 && b < 0); }

__global__ void rds_decode_kernel(const float *rrc, long long ld, int n_streams, int n_blocks, int n, uint8_t *bits_out, int32_t *n_bits_out,
                                  fmrx_rds_event *events, int32_t *n_events_out, int32_t *state) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    int32_t *st = state + (long long)s * FMRX_RDS_STATE_WORDS;
    int block_id = st[W_BLOCK];
    unsigned offset = (unsigned)st[W_OFFSET], start_pos = (unsigned)st[W_START];
    float lonely = __int_as_float(st[W_LONELY]);
    int front_bit = st[W_FRONT], prebit = st[W_PREBIT], nbits = st[W_NBITS];
    unsigned printpos = (unsigned)st[W_PRINTPOS];
    int last_pos = st[W_LASTPOS1] - 1, bad = st[W_BAD];
    uint8_t bits[FMRX_MAX_BITS + 1], diff[27 + FMRX_MAX_BITS + 1];
    for (int i = 0; i < FMRX_MAX_BITS; ++i) bits[i] = (uint8_t)st[W_BITS + i];
    const int nsym = n / SPS;

    for (int b = 0; b < n_blocks; ++b, ++block_id) {
        const float *r = rrc + (long long)s * ld + (long long)b * n;
        if (block_id == 0) {  // :503-517
            float best = fabsf(r[0]);
            for (unsigned i = 1; i < SPS; ++i)
                if (fabsf(r[i]) > best) { best = fabsf(r[i]); offset = i; }
        }
        const float *sym = r + offset;  // sym[k] = r[24k + offset]
        if (block_id == 0) {  // :542-558 (loop index starts at 0, Q10)
            int c0 = 0, c1 = 0;
            for (int j = 0; j < nsym / 4; ++j) {
                const float a0 = sym[SPS * 2 * j], a1 = sym[SPS * (2 * j + 1)], a2 = sym[SPS * (2 * j + 2)];
                if (same_sign(a0, a1)) ++c0;
                else if (same_sign(a1, a2)) ++c1;
            }
            if (c0 > c1) start_pos = 1;
            else if (c1 > c0) start_pos = 0;
        }
        const int want = nsym / 2 - (int)start_pos;  // :560
        for (int i = nbits; i < want; ++i) bits[i] = 0;
        nbits = want;
        if (start_pos == 1 && block_id != 0) {  // :565-572
            const float s0 = sym[0];
            if (lonely > s0) front_bit = 1;
            else if (s0 > lonely) front_bit = 0;
        }
        for (int k = 0; k < nbits; ++k) {  // :574-585
            const unsigned a = 2u * k + start_pos;
            if (a + 1 > (unsigned)nsym - 1) break;
            const float u = sym[SPS * a], v = sym[SPS * (a + 1)];
            if (u > v) bits[k] = 1;
            else if (u < v) bits[k] = 0;
        }
        if (start_pos == 1) {  // :587-592
            for (int i = nbits; i > 0; --i) bits[i] = bits[i - 1];
            bits[0] = (uint8_t)front_bit;
            ++nbits;
            lonely = sym[SPS * (nsym - 1)];
        }
        int off = 0;  // :596-616
        if (block_id == 0) { prebit = bits[0]; off = 1; }
        const int ncarry = block_id != 0 ? 27 : 0;
        for (int g = 0; g < ncarry; ++g) diff[g] = (uint8_t)st[W_CARRY + g];
        const int nd = nbits - off;
        uint8_t *bo = bits_out ? bits_out + ((long long)s * n_blocks + b) * FMRX_MAX_BITS : nullptr;
        for (int t = 0; t < nd; ++t) {
            const int v = prebit ^ bits[t + off];
            diff[ncarry + t] = (uint8_t)v;
            prebit = bits[t + off];
            if (bo) bo[t] = (uint8_t)v;
        }
        prebit = bits[nbits - 1];
        if (n_bits_out) n_bits_out[(long long)s * n_blocks + b] = nd;
        const int total = ncarry + nd;
        fmrx_rds_event *ev = events ? events + ((long long)s * n_blocks + b) * FMRX_MAX_EVENTS : nullptr;
        int nev = 0;
        unsigned pos = 0;
        for (;;) {  // :631-713
            unsigned syn = 0;
            for (int j = 0; j < 26; ++j)
                if (diff[pos + j]) syn ^= kH[j];
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
        for (int g = 0; g < 27; ++g) st[W_CARRY + g] = diff[pos - 1 + g];  // :715-718
        if (n_events_out) n_events_out[(long long)s * n_blocks + b] = nev < FMRX_MAX_EVENTS ? nev : FMRX_MAX_EVENTS;
    }
    st[W_BLOCK] = block_id; st[W_OFFSET] = (int)offset; st[W_START] = (int)start_pos; st[W_LONELY] = __float_as_int(lonely);
    st[W_FRONT] = front_bit; st[W_PREBIT] = prebit; st[W_NBITS] = nbits; st[W_PRINTPOS] = (int)printpos;
    st[W_LASTPOS1] = last_pos + 1; st[W_BAD] = bad;
    for (int i = 0; i < FMRX_MAX_BITS; ++i) st[W_BITS + i] = bits[i];
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
    if (fast < 0) {
        if (j.exact) resample_kernel<true><<<grid, 128, 0, st>>>(d);
        else resample_kernel<false><<<grid, 128, 0, st>>>(d);
        launch_counter() += 1;
    }
    cudaError_t e = cudaGetLastError();
    if (e) return (int)e;
    const int p0 = j.n_ref - j.nzi - 1;
    const bool vec = (j.nzi & 3) == 0 && (p0 & 3) == 0 && p0 >= 0 && p0 + j.nzi <= j.n && (j.ldx & 3) == 0 && (j.n & 3) == 0 &&
                     (((uintptr_t)j.x | (uintptr_t)j.zi) & 15) == 0;
    if (vec) {
        dim3 sg((j.nzi / 4 + 255) / 256, j.n_streams);
        resample_state4_kernel<<<sg, 256, 0, st>>>(d);
    } else {
        dim3 sg((j.nzi + 255) / 256, j.n_streams);
        resample_state_kernel<<<sg, 256, 0, st>>>(d);
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
