// K-quant (Q2_K / Q3_K / Q4_K / Q5_K / Q6_K) and IQ4_NL block math, written as phase functions that run once per
// thread with every cross-thread exchange going through a plain "shared" struct.  On the
// GPU the struct lives in shared memory and phases are separated by __syncthreads(); the
// host test harness (tests/host_emul.cu) runs the same phase functions in a loop over
// thread ids, so the arithmetic is checked on the CPU against the oracle before any GPU time.
//
// Replaces the per-row work of `llama-quantize` launched by the reference at
// ref/src/quantool/methods/llama_cpp/llama_cpp.py:165-178; algorithm = llama.cpp
// ggml-quants.c quantize_row_q4_K_ref / q5_K_ref / q6_K_ref as recorded in SURVEY.md §D.4-§D.5.
// Arithmetic contract: strict fp32 in the source evaluation order, no FMA contraction
// (this translation unit is compiled with --fmad=false), IEEE div/sqrt, fp16 = RNE.
#pragma once
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace qt {
namespace kq {

QT_HD int nearest_int(float fval) {
    // 12582912.f magic constant: round-half-even for |f| < 2^22
    float val = fval + 12582912.f;
#ifdef __CUDA_ARCH__
    int i = __float_as_int(val);
#else
    int i;
    memcpy(&i, &val, sizeof(int));
#endif
    return (i & 0x007fffff) - 0x00400000;
}

QT_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// (float)clamp(nearest_int(v), lo, hi) without leaving the float domain: rounding is monotonic and lo / hi are
// integers, so clamping first and rounding with the same magic constant gives the same value - two min/max (ALU
// pipe) and two adds instead of add + mask + subtract + two integer min/max + I2F.  Valid for |v| < 2^22 like
// nearest_int itself (the candidate searches produce |v| <= nmax + 1).
QT_HD float round_clamp_f(float v, float lo, float hi) {
    v = fminf(fmaxf(v, lo), hi);
    return (v + 12582912.f) - 12582912.f;
}

// ------------------------------------------------------------------------------------
// make_qkx2_quants(n=32, nmax, x, weights=av_x+|x|, rmin, rdelta, nstep, use_mad=false)
// Thread-private.  Instead of keeping L[]/Laux[] arrays it remembers the (iscale, min)
// pair the adopted candidate was generated with; L is regenerated from that pair on demand
// (identical fp32 expression, so identical integers).
// ------------------------------------------------------------------------------------
struct Qkx2Result {
    float scale;      // returned scale
    float the_min;    // *the_min = -min
    float l_iscale;   // L[i] = clamp(nearest_int(l_iscale * (x[i] - l_min)), 0, nmax)
    float l_min;
    int all_zero;     // max == min path: L = 0
};

QT_HD void qkx2_search(const float (&x)[32], int nmax, float rmin, float rdelta, int nstep, Qkx2Result& r) {
    float sum_x2 = 0;
#pragma unroll
    for (int l = 0; l < 32; ++l) sum_x2 += x[l] * x[l];
    const float av_x = sqrtf(sum_x2 / 32);
    const float fmax_ = (float)nmax;

    // weights[i] = av_x + |x[i]| is the same value every time the reference recomputes it: kept in registers
    float wt[32];
    float mn = x[0], mx = x[0];
    wt[0] = av_x + fabsf(x[0]);
    float sum_w = wt[0];
    float sum_x = sum_w * x[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
        wt[i] = av_x + fabsf(x[i]);
        sum_w += wt[i];
        sum_x += wt[i] * x[i];
    }
    if (mn > 0) mn = 0;
    if (mx == mn) {
        r.scale = 0.f; r.the_min = -mn; r.l_iscale = 0.f; r.l_min = mn; r.all_zero = 1;
        return;
    }
    float iscale = nmax / (mx - mn);
    float scale = 1 / iscale;
    float best_error = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float lf = round_clamp_f(iscale * (x[i] - mn), 0.f, fmax_);
        float diff = scale * lf + mn - x[i];
        diff = diff * diff;
        best_error += wt[i] * diff;
    }
    r.l_iscale = iscale; r.l_min = mn; r.all_zero = 0;
    for (int is = 0; is <= nstep; ++is) {
        iscale = (rmin + rdelta * is + nmax) / (mx - mn);
        float lf[32];
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            lf[i] = round_clamp_f(iscale * (x[i] - mn), 0.f, fmax_);
            const float wl = wt[i] * lf[i];
            sum_l += wl;
            sum_l2 += wl * lf[i];
            sum_xl += wl * x[i];
        }
        const float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float cur_error = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float diff = this_scale * lf[i] + this_min - x[i];
                diff = diff * diff;
                cur_error += wt[i] * diff;
            }
            if (cur_error < best_error) {
                r.l_iscale = iscale; r.l_min = mn;   // the pair Laux was generated with
                best_error = cur_error;
                scale = this_scale;
                mn = this_min;                       // C: `min = this_min` feeds later candidates
            }
        }
    }
    r.scale = scale;
    r.the_min = -mn;
}

// ------------------------------------------------------------------------------------
// Q4_K / Q5_K: NSB super-blocks per CTA, 8 threads per super-block (one per 32-elem sub-block)
// ------------------------------------------------------------------------------------
template <int NSB, int OUT_BYTES>
struct K45Shared {
    // x is dead once phase A has copied it to registers, so L/out alias it (48 KB static limit)
    union {
        float x[NSB * 8][33];    // +1 pad: thread t walks row t -> conflict-free
        struct {
            alignas(16) uint8_t out[NSB * OUT_BYTES];
            uint8_t L[NSB][256];
        } o;
    } u;
    float sc[NSB * 8];
    float mn[NSB * 8];
    uint8_t ls[NSB * 8];
    uint8_t lm[NSB * 8];
};

struct K45Thread {
    float x[32];
    Qkx2Result r;
};

template <class S>
QT_HD void k45_phase_a(int t, S& s, K45Thread& th, int nmax, float rmin, float rdelta, int nstep) {
#pragma unroll
    for (int l = 0; l < 32; ++l) th.x[l] = s.u.x[t][l];
    qkx2_search(th.x, nmax, rmin, rdelta, nstep, th.r);
    s.sc[t] = th.r.scale;
    s.mn[t] = th.r.the_min;
}

template <class S, int OUT_BYTES>
QT_HD void k45_phase_b(int t, S& s, const K45Thread& th, int nmax) {
    const int sb = t >> 3, j = t & 7;
    float max_scale = 0, max_min = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float scale = s.sc[sb * 8 + k];
        if (scale > max_scale) max_scale = scale;
        const float m = s.mn[sb * 8 + k];
        if (m > max_min) max_min = m;
    }
    const float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    const float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    uint8_t ls = (uint8_t)nearest_int(inv_scale * th.r.scale);
    uint8_t lm = (uint8_t)nearest_int(inv_min * th.r.the_min);
    ls = ls < 63 ? ls : 63;
    lm = lm < 63 ? lm : 63;
    s.ls[t] = ls;
    s.lm[t] = lm;
    const __half dh = __float2half_rn(max_scale / 63.f);
    const __half mh = __float2half_rn(max_min / 63.f);
    if (j == 0) {
        uint8_t* o = s.u.o.out + sb * OUT_BYTES;
        const unsigned short db = __half_as_ushort(dh), mb = __half_as_ushort(mh);
        o[0] = (uint8_t)(db & 0xff); o[1] = (uint8_t)(db >> 8);
        o[2] = (uint8_t)(mb & 0xff); o[3] = (uint8_t)(mb >> 8);
    }
    // get_scale_min_k4 returns exactly (ls, lm) after the 6-bit packing round trip
    const float d = __half2float(dh) * ls;
    uint8_t* L = &s.u.o.L[sb][32 * j];
    if (d != 0.f) {
        const float dm = __half2float(mh) * lm;
#pragma unroll
        for (int ii = 0; ii < 32; ++ii) L[ii] = (uint8_t)clampi(nearest_int((th.x[ii] + dm) / d), 0, nmax);
    } else if (th.r.all_zero) {
#pragma unroll
        for (int ii = 0; ii < 32; ++ii) L[ii] = 0;
    } else {
#pragma unroll
        for (int ii = 0; ii < 32; ++ii)
            L[ii] = (uint8_t)clampi(nearest_int(th.r.l_iscale * (th.x[ii] - th.r.l_min)), 0, nmax);
    }
}

// scales[12] packing, common to Q4_K and Q5_K (block offset 4)
template <class S, int OUT_BYTES>
QT_HD void k45_pack_scales(int t, S& s) {
    const int sb = t >> 3, j = t & 7;
    uint8_t* sc = s.u.o.out + sb * OUT_BYTES + 4;
    const uint8_t ls = s.ls[t], lm = s.lm[t];
    if (j < 4) {
        const uint8_t ls4 = s.ls[t + 4], lm4 = s.lm[t + 4];
        sc[j] = (uint8_t)(ls | ((ls4 >> 4) << 6));
        sc[j + 4] = (uint8_t)(lm | ((lm4 >> 4) << 6));
    } else {
        sc[j + 4] = (uint8_t)((ls & 0xF) | ((lm & 0xF) << 4));
    }
}

template <class S>
QT_HD void q4k_phase_c(int t, S& s) {
    k45_pack_scales<S, 144>(t, s);
    const int sb = t >> 3, j = t & 7;
    // qs byte b (0..127): chunk c=b/32, l=b%32 -> L[64c+l] | L[64c+32+l]<<4 ; thread j: bytes 16j..16j+15
    uint8_t* q = s.u.o.out + sb * 144 + 16 + 16 * j;
    const int c = j >> 1, l0 = (j & 1) * 16;
    const uint8_t* L = s.u.o.L[sb];
#pragma unroll
    for (int i = 0; i < 16; ++i) q[i] = (uint8_t)(L[64 * c + l0 + i] | (L[64 * c + 32 + l0 + i] << 4));
}

template <class S>
QT_HD void q5k_phase_c(int t, S& s) {
    k45_pack_scales<S, 176>(t, s);
    const int sb = t >> 3, j = t & 7;
    const uint8_t* L = s.u.o.L[sb];
    uint8_t* qh = s.u.o.out + sb * 176 + 16;
    uint8_t* ql = s.u.o.out + sb * 176 + 48 + 16 * j;
    const int c = j >> 1, l0 = (j & 1) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int l1 = L[64 * c + l0 + i], l2 = L[64 * c + 32 + l0 + i];
        ql[i] = (uint8_t)((l1 & 0xF) | ((l2 & 0xF) << 4));
    }
    // qh[jj] for jj = 4j..4j+3: bit 2c <- L[64c+jj] > 15, bit 2c+1 <- L[64c+32+jj] > 15
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int jj = 4 * j + k;
        uint8_t h = 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            h |= (uint8_t)((L[64 * cc + jj] > 15 ? 1 : 0) << (2 * cc));
            h |= (uint8_t)((L[64 * cc + 32 + jj] > 15 ? 1 : 0) << (2 * cc + 1));
        }
        qh[jj] = h;
    }
}

// ------------------------------------------------------------------------------------
// Q6_K: 16 threads per super-block (one per 16-elem sub-block)
// ------------------------------------------------------------------------------------
template <int NSB>
struct K6Shared {
    float x[NSB * 16][17];
    float sc[NSB * 16];
    uint8_t L[NSB][256];
    alignas(16) uint8_t out[NSB * 210];
};

struct K6Thread {
    float x[16];
    float scale;
    float l_iscale;  // L[i] = 32 + clamp(nearest_int(l_iscale * x[i]), -32, 31)
    int all_zero;
};

// make_qx_quants(n=16, nmax=32, rmse_type=1, qw=NULL)
QT_HD void qx_search(K6Thread& th) {
    const int nmax = 32;
    float mx = 0, amax = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ax = fabsf(th.x[i]);
        if (ax > amax) { amax = ax; mx = th.x[i]; }
    }
    if (amax < 1e-15f) {
        th.scale = 0.f; th.l_iscale = 0.f; th.all_zero = 1;
        return;
    }
    th.all_zero = 0;
    // w = x*x and w*x are the same values in every candidate (C evaluates w*x[i]*l as (w*x[i])*l): kept in registers
    float w[16], wx[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { w[i] = th.x[i] * th.x[i]; wx[i] = w[i] * th.x[i]; }
    float iscale = -nmax / mx;
    float sumlx = 0, suml2 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float l = round_clamp_f(iscale * th.x[i], -32.f, 31.f);
        sumlx += wx[i] * l;
        suml2 += w[i] * l * l;
    }
    float scale = suml2 ? sumlx / suml2 : 0.0f;
    float best = scale * sumlx;
    th.l_iscale = iscale;
    for (int is = -9; is <= 9; ++is) {
        if (is == 0) continue;
        iscale = -(nmax + 0.1f * is) / mx;
        sumlx = suml2 = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float l = round_clamp_f(iscale * th.x[i], -32.f, 31.f);
            sumlx += wx[i] * l;
            suml2 += w[i] * l * l;
        }
        if (suml2 > 0 && sumlx * sumlx > best * suml2) {
            th.l_iscale = iscale;
            scale = sumlx / suml2;
            best = scale * sumlx;
        }
    }
    th.scale = scale;
}

template <class S>
QT_HD void q6k_phase_a(int t, S& s, K6Thread& th) {
#pragma unroll
    for (int l = 0; l < 16; ++l) th.x[l] = s.x[t][l];
    qx_search(th);
    s.sc[t] = th.scale;
}

template <class S>
QT_HD void q6k_phase_b(int t, S& s, const K6Thread& th) {
    const int sb = t >> 4, j = t & 15;
    float max_scale = 0, max_abs_scale = 0;
#pragma unroll
    for (int ib = 0; ib < 16; ++ib) {
        const float scale = s.sc[sb * 16 + ib];
        const float a = fabsf(scale);
        if (a > max_abs_scale) { max_abs_scale = a; max_scale = scale; }
    }
    uint8_t* o = s.out + sb * 210;
    uint8_t* L = &s.L[sb][16 * j];
    if (max_abs_scale < 1e-15f) {
        // memset(block, 0); d = fp16(0)
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = 0;
        o[192 + j] = 0;
        if (j == 0) { o[208] = 0; o[209] = 0; }
        return;
    }
    const float iscale = -128.f / max_scale;
    const __half dh = __float2half_rn(1 / iscale);
    int q = nearest_int(iscale * th.scale);
    q = q < 127 ? q : 127;
    const int8_t q8 = (int8_t)q;
    o[192 + j] = (uint8_t)q8;
    if (j == 0) {
        const unsigned short db = __half_as_ushort(dh);
        o[208] = (uint8_t)(db & 0xff); o[209] = (uint8_t)(db >> 8);
    }
    const float d = __half2float(dh) * q8;
    if (d != 0.f) {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = (uint8_t)(clampi(nearest_int(th.x[ii] / d), -32, 31) + 32);
    } else if (th.all_zero) {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = 0;
    } else {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii)
            L[ii] = (uint8_t)(32 + clampi(nearest_int(th.l_iscale * th.x[ii]), -32, 31));
    }
}

// returns 1 when the super-block is the all-zero special case (whole block memset to 0)
template <class S>
QT_HD void q6k_phase_c(int t, S& s) {
    const int sb = t >> 4, j = t & 15;
    const uint8_t* L = s.L[sb];
    uint8_t* o = s.out + sb * 210;
    // ql: 128 bytes; thread j writes bytes 8j..8j+7.  byte b: half h=b/64, r=b%64;
    //   r<32 : L[128h + r] & 15 | (L[128h + r + 64] & 15) << 4
    //   r>=32: L[128h + r] & 15 | (L[128h + r + 64] & 15) << 4    (r-32+32 = r; +96 = r+64)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int b = 8 * j + i, h = b >> 6, r = b & 63;
        o[b] = (uint8_t)((L[128 * h + r] & 0xF) | ((L[128 * h + r + 64] & 0xF) << 4));
    }
    // qh: 64 bytes; thread j writes bytes 4j..4j+3.  byte b: h=b/32, l=b%32
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int b = 4 * j + i, h = b >> 5, l = b & 31;
        const uint8_t* Lh = L + 128 * h;
        o[128 + b] = (uint8_t)((Lh[l] >> 4) | ((Lh[l + 32] >> 4) << 2) | ((Lh[l + 64] >> 4) << 4) |
                               ((Lh[l + 96] >> 4) << 6));
    }
}

// ------------------------------------------------------------------------------------
// Q2_K / Q3_K: 16 threads per super-block (one per 16-elem sub-block), like Q6_K.
//   Q2_K = llama.cpp quantize_row_q2_K_ref: make_qkx2_quants(16, 3, x, weights=|x|, -0.5, 0.1, 15,
//          use_mad=true), 4-bit scale and 4-bit min per sub-block against fp16 d / dmin.
//   Q3_K = quantize_row_q3_K_ref: make_q3_quants(16, 4, x, do_rmse=true) (coordinate descent,
//          sequential per sub-block), 6-bit signed sub-scales against fp16 d.
// ------------------------------------------------------------------------------------
template <int NSB, int OUT_BYTES>
struct K23Shared {
    float x[NSB * 16][17];
    float sc[NSB * 16];
    float mn[NSB * 16];
    uint8_t l6[NSB * 16];
    uint8_t L[NSB][256];
    alignas(16) uint8_t out[NSB * OUT_BYTES];
};

struct K2Thread {
    float x[16];
    Qkx2Result r;
};

// make_qkx2_quants(n=16, nmax=3, weights=|x|, rmin=-0.5, rdelta=0.1, nstep=15, use_mad=true)
QT_HD void qkx2_search_q2(const float (&x)[16], Qkx2Result& r) {
    const int nmax = 3, nstep = 15;
    const float rmin = -0.5f, rdelta = 0.1f;
    float mn = x[0], mx = x[0];
    float sum_w = fabsf(x[0]);
    float sum_x = sum_w * x[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
        const float w = fabsf(x[i]);
        sum_w += w;
        sum_x += w * x[i];
    }
    if (mn > 0) mn = 0;
    if (mx == mn) {
        r.scale = 0.f; r.the_min = -mn; r.l_iscale = 0.f; r.l_min = mn; r.all_zero = 1;
        return;
    }
    float iscale = nmax / (mx - mn);
    float scale = 1 / iscale;
    float best_mad = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float l = round_clamp_f(iscale * (x[i] - mn), 0.f, 3.f);
        const float diff = fabsf(scale * l + mn - x[i]);
        best_mad += fabsf(x[i]) * diff;
    }
    r.l_iscale = iscale; r.l_min = mn; r.all_zero = 0;
    for (int is = 0; is <= nstep; ++is) {
        iscale = (rmin + rdelta * is + nmax) / (mx - mn);
        float lf[16];
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            lf[i] = round_clamp_f(iscale * (x[i] - mn), 0.f, 3.f);
            const float wl = fabsf(x[i]) * lf[i];
            sum_l += wl;
            sum_l2 += wl * lf[i];
            sum_xl += wl * x[i];
        }
        const float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float mad = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float diff = fabsf(this_scale * lf[i] + this_min - x[i]);
                mad += fabsf(x[i]) * diff;
            }
            if (mad < best_mad) {
                r.l_iscale = iscale; r.l_min = mn;
                best_mad = mad;
                scale = this_scale;
                mn = this_min;
            }
        }
    }
    r.scale = scale;
    r.the_min = -mn;
}

template <class S>
QT_HD void q2k_phase_a(int t, S& s, K2Thread& th) {
#pragma unroll
    for (int l = 0; l < 16; ++l) th.x[l] = s.x[t][l];
    qkx2_search_q2(th.x, th.r);
    s.sc[t] = th.r.scale;
    s.mn[t] = th.r.the_min;
}

template <class S>
QT_HD void q2k_phase_b(int t, S& s, const K2Thread& th) {
    const int sb = t >> 4, j = t & 15;
    float max_scale = 0, max_min = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float scale = s.sc[sb * 16 + k];
        if (scale > max_scale) max_scale = scale;
        const float m = s.mn[sb * 16 + k];
        if (m > max_min) max_min = m;
    }
    const float q4scale = 15.f;
    uint8_t byte = 0;
    __half dh = __float2half_rn(0.f), mh = __float2half_rn(0.f);
    if (max_scale > 0) {
        const float iscale = q4scale / max_scale;
        byte = (uint8_t)nearest_int(iscale * th.r.scale);
        dh = __float2half_rn(max_scale / q4scale);
    }
    if (max_min > 0) {
        const float iscale = q4scale / max_min;
        const int l = nearest_int(iscale * th.r.the_min);
        byte |= (uint8_t)(l << 4);
        mh = __float2half_rn(max_min / q4scale);
    }
    uint8_t* o = s.out + sb * 84;
    o[j] = byte;
    if (j == 0) {
        const unsigned short db = __half_as_ushort(dh), mb = __half_as_ushort(mh);
        o[80] = (uint8_t)(db & 0xff); o[81] = (uint8_t)(db >> 8);
        o[82] = (uint8_t)(mb & 0xff); o[83] = (uint8_t)(mb >> 8);
    }
    const float d = __half2float(dh) * (byte & 0xF);
    uint8_t* L = &s.L[sb][16 * j];
    if (d != 0.f) {
        const float dm = __half2float(mh) * (byte >> 4);
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = (uint8_t)clampi(nearest_int((th.x[ii] + dm) / d), 0, 3);
    } else if (th.r.all_zero) {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = 0;
    } else {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii)
            L[ii] = (uint8_t)clampi(nearest_int(th.r.l_iscale * (th.x[ii] - th.r.l_min)), 0, 3);
    }
}

// qs byte b (0..63): half h=b/32, l=b%32 -> 2-bit codes of elements 128h + l + {0,32,64,96}
QT_HD uint8_t pack2_byte(const uint8_t* L, int b) {
    const uint8_t* Lh = L + 128 * (b >> 5) + (b & 31);
    return (uint8_t)((Lh[0] & 3) | ((Lh[32] & 3) << 2) | ((Lh[64] & 3) << 4) | ((Lh[96] & 3) << 6));
}

template <class S>
QT_HD void q2k_phase_c(int t, S& s) {
    const int sb = t >> 4, j = t & 15;
    uint8_t* q = s.out + sb * 84 + 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) q[4 * j + i] = pack2_byte(s.L[sb], 4 * j + i);
}

struct K3Thread {
    float x[16];
    float scale;
    int L[16];       // make_q3_quants result, already offset by +4 (0 for an all-zero sub-block)
};

// make_q3_quants(n=16, nmax=4, x, L, do_rmse=true)
QT_HD void q3_search(K3Thread& th) {
    const int nmax = 4;
    float mx = 0, amax = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float ax = fabsf(th.x[i]);
        if (ax > amax) { amax = ax; mx = th.x[i]; }
    }
    if (amax < 1e-15f) {
#pragma unroll
        for (int i = 0; i < 16; ++i) th.L[i] = 0;
        th.scale = 0.f;
        return;
    }
    const float iscale = -nmax / mx;
    float sumlx = 0, suml2 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int l = clampi(nearest_int(iscale * th.x[i]), -nmax, nmax - 1);
        th.L[i] = l;
        const float w = th.x[i] * th.x[i];
        sumlx += w * th.x[i] * l;
        suml2 += w * l * l;
    }
#pragma unroll 1
    for (int itry = 0; itry < 5; ++itry) {
        int n_changed = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float w = th.x[i] * th.x[i];
            float slx = sumlx - w * th.x[i] * th.L[i];
            if (slx > 0) {
                float sl2 = suml2 - w * th.L[i] * th.L[i];
                int new_l = nearest_int(th.x[i] * sl2 / slx);
                new_l = clampi(new_l, -nmax, nmax - 1);
                if (new_l != th.L[i]) {
                    slx += w * th.x[i] * new_l;
                    sl2 += w * new_l * new_l;
                    if (sl2 > 0 && slx * slx * suml2 > sumlx * sumlx * sl2) {
                        th.L[i] = new_l; sumlx = slx; suml2 = sl2;
                        ++n_changed;
                    }
                }
            }
        }
        if (!n_changed) break;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) th.L[i] += nmax;
    th.scale = sumlx / suml2;
}

template <class S>
QT_HD void q3k_phase_a(int t, S& s, K3Thread& th) {
#pragma unroll
    for (int l = 0; l < 16; ++l) th.x[l] = s.x[t][l];
    q3_search(th);
    s.sc[t] = th.scale;
}

template <class S>
QT_HD void q3k_phase_b(int t, S& s, const K3Thread& th) {
    const int sb = t >> 4, j = t & 15;
    float max_scale = 0, amax = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float v = s.sc[sb * 16 + k];
        const float a = fabsf(v);
        if (a > amax) { amax = a; max_scale = v; }
    }
    uint8_t* o = s.out + sb * 110;
    int l6 = 0;          // stored 6-bit value; the packed bytes stay zero when max_scale == 0
    __half dh = __float2half_rn(0.f);
    if (max_scale != 0.f) {
        const float iscale = -32.f / max_scale;
        const int8_t l = (int8_t)nearest_int(iscale * th.scale);
        l6 = clampi((int)l, -32, 31) + 32;
        dh = __float2half_rn(1 / iscale);
    }
    s.l6[t] = (uint8_t)l6;
    if (j == 0) {
        const unsigned short db = __half_as_ushort(dh);
        o[108] = (uint8_t)(db & 0xff); o[109] = (uint8_t)(db >> 8);
    }
    const int sc = l6 - 32;   // what the C re-reads from the packed bytes (0 - 32 when they are zero)
    const float d = __half2float(dh) * sc;
    uint8_t* L = &s.L[sb][16 * j];
    if (d != 0.f) {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = (uint8_t)(clampi(nearest_int(th.x[ii] / d), -4, 3) + 4);
    } else {
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) L[ii] = (uint8_t)th.L[ii];
    }
}

template <class S>
QT_HD void q3k_phase_c(int t, S& s) {
    const int sb = t >> 4, j = t & 15;
    const uint8_t* L = s.L[sb];
    uint8_t* o = s.out + sb * 110;
    // hmask byte m (0..31): bit b <- code of element 32b + m has its high bit set
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = 2 * j + i;
        uint8_t h = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) h |= (uint8_t)((L[32 * b + m] > 3 ? 1 : 0) << b);
        o[m] = h;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) o[32 + 4 * j + i] = pack2_byte(L, 4 * j + i);
    // 6-bit scales: bytes 0..7 = low nibbles of sub-blocks b and b+8; bytes 8..11 = the 2 high bits
    const uint8_t* l6 = &s.l6[sb * 16];
    if (j < 8) {
        o[96 + j] = (uint8_t)((l6[j] & 0xF) | ((l6[j + 8] & 0xF) << 4));
    } else if (j < 12) {
        const int k = j - 8;
        o[96 + j] = (uint8_t)((l6[k] >> 4) | ((l6[k + 4] >> 4) << 2) | ((l6[k + 8] >> 4) << 4) | ((l6[k + 12] >> 4) << 6));
    }
}

// ------------------------------------------------------------------------------------
// IQ4_NL (llama.cpp's fallback for Q2_K / Q3_K rows that are not a multiple of 256): one thread per 32-element
// block.  quantize_row_iq4_nl_impl(32, 32, quant_weights = NULL, ntry = 7): weights x^2, 1 + 15 candidate scales.
// ------------------------------------------------------------------------------------
// kvalues_iq4nl.  The packer indexes the table with data-dependent indices: a `switch` became divergent branches
// (one thread per block, every lane somewhere else: 155 ms for 235 M elements), so the callers pass a 16-entry
// table - shared memory on the device (16 distinct banks: conflict-free), a static array on the host.
#define QT_IQ4NL_VALUES {-127.f, -104.f, -83.f, -65.f, -49.f, -35.f, -22.f, -10.f, 1.f, 13.f, 25.f, 38.f, 53.f, 69.f, 89.f, 113.f}

QT_HD float iq4nl_value(int i) {
    // immediates, for compile-time or warp-uniform indices (dequantize)
    switch (i) {
        case 0: return -127.f; case 1: return -104.f; case 2: return -83.f; case 3: return -65.f;
        case 4: return -49.f; case 5: return -35.f; case 6: return -22.f; case 7: return -10.f;
        case 8: return 1.f; case 9: return 13.f; case 10: return 25.f; case 11: return 38.f;
        case 12: return 53.f; case 13: return 69.f; case 14: return 89.f; default: return 113.f;
    }
}

// best_index_int8(16, kvalues_iq4nl, x): nearest table value, ties to the upper one.  Branch-free form of the
// reference's bisection: with 16 entries it ends after at most 4 halvings, and a halving of an interval of width 1
// is a no-op (mav == ml, x >= tbl[ml] holds), so four unconditional rounds give the same (ml, mu).
// Returns the index and (through `q`) the table value at that index.  The first two bisection rounds compare
// against immediates (the packer was bound by shared-memory look-ups: 7 per element and candidate, now 4).
QT_HD int iq4nl_best_index_q(const float* tbl, float x, float& q) {
    // round 1: mav = 7 (-10); round 2: mav = 3 (-65) or 11 (38)
    const bool b1 = x < -10.f;
    const bool b2 = x < (b1 ? -65.f : 38.f);
    int ml = b1 ? (b2 ? 0 : 3) : (b2 ? 7 : 11);
    int mu = b1 ? (b2 ? 3 : 7) : (b2 ? 11 : 15);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int mav = (ml + mu) >> 1;
        const bool below = x < tbl[mav];
        mu = below ? mav : mu;
        ml = below ? ml : mav;
    }
    const int lo = mu > 0 ? mu - 1 : 0;
    const float a = tbl[lo], b = tbl[mu];
    const bool take_lo = (x - a < b - x);
    int r = take_lo ? lo : mu;
    q = take_lo ? a : b;
    if (x >= 113.f) { r = 15; q = 113.f; }
    if (x <= -127.f) { r = 0; q = -127.f; }
    return r;
}
QT_HD int iq4nl_best_index(const float* tbl, float x) {
    float q;
    return iq4nl_best_index_q(tbl, x, q);
}

QT_HD void iq4nl_sums(const float* tbl, const float (&x)[32], float id, float& sumqx, float& sumq2) {
    sumqx = 0.f; sumq2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float q;
        iq4nl_best_index_q(tbl, id * x[j], q);
        const float w = x[j] * x[j];
        sumqx += w * q * x[j];
        sumq2 += w * q * q;
    }
}

// out: 18 bytes (fp16 d, 16 bytes of nibbles: element j low, j + 16 high)
QT_HD void iq4nl_block(const float* tbl, const float (&x)[32], uint8_t* out) {
    float amax = 0.f, mx = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float ax = fabsf(x[j]);
        if (ax > amax) { amax = ax; mx = x[j]; }
    }
    float scale = 0.f;
    if (!(amax < 1e-15f)) {
        float d = -mx / -127.f;
        float sumqx, sumq2;
        iq4nl_sums(tbl, x, 1 / d, sumqx, sumq2);
        d = sumqx / sumq2;
        float best = d * sumqx;
#pragma unroll 1
        for (int itry = -7; itry <= 7; ++itry) {
            const float id = (itry + -127.f) / mx;
            iq4nl_sums(tbl, x, id, sumqx, sumq2);
            if (sumq2 > 0 && sumqx * sumqx > best * sumq2) {
                d = sumqx / sumq2; best = d * sumqx;
            }
        }
        scale = d;
    }
    const unsigned short db = __half_as_ushort(__float2half_rn(scale));
    out[0] = (uint8_t)(db & 0xff); out[1] = (uint8_t)(db >> 8);
    const float id = scale ? 1 / scale : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j)
        out[2 + j] = (uint8_t)(iq4nl_best_index(tbl, id * x[j]) | (iq4nl_best_index(tbl, id * x[16 + j]) << 4));
}

}  // namespace kq
}  // namespace qt
