/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * CPU restatement of llama.cpp's reference block quantizers, the arithmetic that
 * `llama-quantize` runs when the reference shells out to it at
 * ref/src/quantool/methods/llama_cpp/llama_cpp.py:165-178 (`_quantize_gguf`).
 *
 * llama.cpp is an UN-VENDORED, UNPINNED dependency of the reference
 * (ref/pyproject.toml:46-48 `llama-cpp-scripts @ git+...llama.cpp@master`,
 * ref/Dockerfile:16-21 `git clone --depth=1`).  No llama.cpp source or binary exists on
 * this box, so this file restates the published algorithm of ggml/src/ggml-quants.c
 * (`quantize_row_<type>_ref`, `make_qkx2_quants`, `make_qx_quants`, `nearest_int`) as
 * recorded in SURVEY.md §D.1-§D.5.
 *
 * Parity pins (tests/test_oracle_cpu.py):
 *   - Q8_0 / Q4_0 / Q4_1 / Q5_0 / Q5_1 packed bytes == gguf-py `gguf.quants.quantize`
 *     (GGUFPY/quants.py:220-239, 291-311, 378-393; "bit-exact same results as reference
 *     implementation in ggml-quants.c") and the sha256 KATs in SURVEY.md §8c.
 *   - IQ4_NL / Q2_K / Q3_K / Q4_K / Q5_K / Q6_K: PARITY UNPINNED against llama.cpp itself (no quantize twin on
 *     disk); pinned only by (a) an independent numpy restatement (oracle/ggml_quants_np.py)
 *     agreeing byte-for-byte and (b) gguf-py's dequantizers (GGUFPY/quants.py:475-572)
 *     reading the packed layout back to within the format's error, and (c) llama.cpp's own
 *     acceptance limits for the round-trip error of each type (tests/test-quantize-fns.cpp, restated).
 *
 * Arithmetic contract: strict IEEE fp32, evaluation order as written, NO fused
 * multiply-add (build with -ffp-contract=off), fp16 conversion = round-to-nearest-even.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fno-fast-math -pthread).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define QK 32
#define QK_K 256
#define GROUP_MAX_EPS 1e-15f

#define MINI(a, b) ((a) < (b) ? (a) : (b))
#define MAXI(a, b) ((a) > (b) ? (a) : (b))

/* ---- fp16 <-> fp32, IEEE round-to-nearest-even, subnormals honoured -------------- */
static uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x007fffffu;
    int32_t exp = (int32_t)((x >> 23) & 0xff);
    if (exp == 0xff) return (uint16_t)(sign | 0x7c00u | (mant ? 0x200u | (mant >> 13) : 0));
    int32_t e = exp - 127 + 15;
    if (e >= 0x1f) return (uint16_t)(sign | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)sign;
        mant |= 0x00800000u;
        uint32_t shift = (uint32_t)(14 - e);
        uint32_t half = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half & 1))) half++;
        return (uint16_t)(sign | half);
    }
    uint32_t half = ((uint32_t)e << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
    return (uint16_t)(sign | half);
}

static float f16_to_f32(uint16_t h) {
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t mant = h & 0x3ffu;
    uint32_t x;
    if (exp == 0) {
        if (mant == 0) {
            x = sign;
        } else {
            int e = -1;
            do { e++; mant <<= 1; } while (!(mant & 0x400u));
            mant &= 0x3ffu;
            x = sign | ((uint32_t)(127 - 15 - e) << 23) | (mant << 13);
        }
    } else if (exp == 0x1f) {
        x = sign | 0x7f800000u | (mant << 13);
    } else {
        x = sign | ((exp - 15 + 127) << 23) | (mant << 13);
    }
    float f;
    memcpy(&f, &x, 4);
    return f;
}

/* nearest_int via the 12582912.f magic constant (round-half-even, |f| < 2^22) */
static inline int nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i;
    memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* ---- block structs (little-endian, packed; GGUFPY/constants.py:4231-4267) ---------- */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } block_q4_0;                 /* 18 B */
typedef struct { uint16_t d; uint16_t m; uint8_t qs[16]; } block_q4_1;     /* 20 B */
typedef struct { uint16_t d; uint8_t qh[4]; uint8_t qs[16]; } block_q5_0;  /* 22 B */
typedef struct { uint16_t d; uint16_t m; uint8_t qh[4]; uint8_t qs[16]; } block_q5_1; /* 24 B */
typedef struct { uint16_t d; int8_t qs[32]; } block_q8_0;                  /* 34 B */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qs[128]; } block_q4_K; /* 144 B */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t qs[128]; } block_q5_K; /* 176 B */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } block_q6_K; /* 210 B */
typedef struct { uint8_t scales[16]; uint8_t qs[64]; uint16_t d; uint16_t dmin; } block_q2_K; /* 84 B */
typedef struct { uint8_t hmask[32]; uint8_t qs[64]; uint8_t scales[12]; uint16_t d; } block_q3_K; /* 110 B */
typedef struct { uint16_t d; uint8_t qs[16]; } block_iq4_nl;               /* 18 B */
#pragma pack(pop)

/* ---- D.1 Q8_0 ------------------------------------------------------------------- */
static void row_q8_0(const float *x, block_q8_0 *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int j = 0; j < QK; j++) {
            const float v = x[i * QK + j];
            amax = amax > fabsf(v) ? amax : fabsf(v);
        }
        const float d = amax / 127;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = f32_to_f16(d);
        for (int j = 0; j < QK; ++j) {
            const float x0 = x[i * QK + j] * id;
            y[i].qs[j] = (int8_t)roundf(x0);
        }
    }
}

/* ---- D.2 Q4_0 ------------------------------------------------------------------- */
static void row_q4_0(const float *x, block_q4_0 *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < QK; j++) {
            const float v = x[i * QK + j];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -8;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = f32_to_f16(d);
        for (int j = 0; j < QK / 2; ++j) {
            const float x0 = x[i * QK + 0 + j] * id;
            const float x1 = x[i * QK + QK / 2 + j] * id;
            const uint8_t xi0 = (uint8_t)MINI(15, (int8_t)(x0 + 8.5f));
            const uint8_t xi1 = (uint8_t)MINI(15, (int8_t)(x1 + 8.5f));
            y[i].qs[j] = xi0 | (uint8_t)(xi1 << 4);
        }
    }
}

static void row_q4_1(const float *x, block_q4_1 *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        float min = 3.402823466e+38f, max = -3.402823466e+38f;
        for (int j = 0; j < QK; j++) {
            const float v = x[i * QK + j];
            if (v < min) min = v;
            if (v > max) max = v;
        }
        const float d = (max - min) / ((1 << 4) - 1);
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = f32_to_f16(d);
        y[i].m = f32_to_f16(min);
        for (int j = 0; j < QK / 2; ++j) {
            const float x0 = (x[i * QK + 0 + j] - min) * id;
            const float x1 = (x[i * QK + QK / 2 + j] - min) * id;
            const uint8_t xi0 = (uint8_t)MINI(15, (int8_t)(x0 + 0.5f));
            const uint8_t xi1 = (uint8_t)MINI(15, (int8_t)(x1 + 0.5f));
            y[i].qs[j] = xi0 | (uint8_t)(xi1 << 4);
        }
    }
}

/* ---- D.3 Q5_0 / Q5_1 ------------------------------------------------------------ */
static void row_q5_0(const float *x, block_q5_0 *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < QK; j++) {
            const float v = x[i * QK + j];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -16;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = f32_to_f16(d);
        uint32_t qh = 0;
        for (int j = 0; j < QK / 2; ++j) {
            const float x0 = x[i * QK + 0 + j] * id;
            const float x1 = x[i * QK + QK / 2 + j] * id;
            const uint8_t xi0 = (uint8_t)MINI(31, (int8_t)(x0 + 16.5f));
            const uint8_t xi1 = (uint8_t)MINI(31, (int8_t)(x1 + 16.5f));
            y[i].qs[j] = (xi0 & 0x0F) | (uint8_t)((xi1 & 0x0F) << 4);
            qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
            qh |= ((xi1 & 0x10u) >> 4) << (j + QK / 2);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}

static void row_q5_1(const float *x, block_q5_1 *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        float min = 3.402823466e+38f, max = -3.402823466e+38f;
        for (int j = 0; j < QK; j++) {
            const float v = x[i * QK + j];
            if (v < min) min = v;
            if (v > max) max = v;
        }
        const float d = (max - min) / ((1 << 5) - 1);
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = f32_to_f16(d);
        y[i].m = f32_to_f16(min);
        uint32_t qh = 0;
        for (int j = 0; j < QK / 2; ++j) {
            const float x0 = (x[i * QK + 0 + j] - min) * id;
            const float x1 = (x[i * QK + QK / 2 + j] - min) * id;
            const uint8_t xi0 = (uint8_t)(x0 + 0.5f);
            const uint8_t xi1 = (uint8_t)(x1 + 0.5f);
            y[i].qs[j] = (xi0 & 0x0F) | (uint8_t)((xi1 & 0x0F) << 4);
            qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
            qh |= ((xi1 & 0x10u) >> 4) << (j + QK / 2);
        }
        memcpy(y[i].qh, &qh, 4);
    }
}

/* ---- D.4 make_qkx2_quants (shared by Q2_K / Q4_K / Q5_K) ------------------------------- */
static float make_qkx2_quants(int n, int nmax, const float *x, const float *weights, uint8_t *L,
                              float *the_min, uint8_t *Laux, float rmin, float rdelta, int nstep,
                              int use_mad) {
    float min = x[0];
    float max = x[0];
    float sum_w = weights[0];
    float sum_x = sum_w * x[0];
    for (int i = 1; i < n; ++i) {
        if (x[i] < min) min = x[i];
        if (x[i] > max) max = x[i];
        float w = weights[i];
        sum_w += w;
        sum_x += w * x[i];
    }
    if (min > 0) min = 0;
    if (max == min) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        *the_min = -min;
        return 0.f;
    }
    float iscale = nmax / (max - min);
    float scale = 1 / iscale;
    float best_error = 0;
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * (x[i] - min));
        L[i] = (uint8_t)MAXI(0, MINI(nmax, l));
        float diff = scale * L[i] + min - x[i];
        diff = use_mad ? fabsf(diff) : diff * diff;
        float w = weights[i];
        best_error += w * diff;
    }
    if (nstep < 1) {
        *the_min = -min;
        return scale;
    }
    for (int is = 0; is <= nstep; ++is) {
        iscale = (rmin + rdelta * is + nmax) / (max - min);
        float sum_l = 0, sum_l2 = 0, sum_xl = 0;
        for (int i = 0; i < n; ++i) {
            int l = nearest_int(iscale * (x[i] - min));
            l = MAXI(0, MINI(nmax, l));
            Laux[i] = (uint8_t)l;
            float w = weights[i];
            sum_l += w * l;
            sum_l2 += w * l * l;
            sum_xl += w * l * x[i];
        }
        float D = sum_w * sum_l2 - sum_l * sum_l;
        if (D > 0) {
            float this_scale = (sum_w * sum_xl - sum_x * sum_l) / D;
            float this_min = (sum_l2 * sum_x - sum_l * sum_xl) / D;
            if (this_min > 0) {
                this_min = 0;
                this_scale = sum_xl / sum_l2;
            }
            float cur_error = 0;
            for (int i = 0; i < n; ++i) {
                float diff = this_scale * Laux[i] + this_min - x[i];
                diff = use_mad ? fabsf(diff) : diff * diff;
                float w = weights[i];
                cur_error += w * diff;
            }
            if (cur_error < best_error) {
                for (int i = 0; i < n; ++i) L[i] = Laux[i];
                best_error = cur_error;
                scale = this_scale;
                min = this_min;
            }
        }
    }
    *the_min = -min;
    return scale;
}

static inline void get_scale_min_k4(int j, const uint8_t *q, uint8_t *d, uint8_t *m) {
    if (j < 4) {
        *d = q[j] & 63;
        *m = q[j + 4] & 63;
    } else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4);
    }
}

/* shared front half of Q4_K / Q5_K: sub-block search, 6-bit scale/min packing, fp16 d/dmin */
static void k45_scales(const float *x, int nmax, float rmin, float rdelta, int nstep, uint8_t *L,
                       uint8_t *scales_out, uint16_t *d_out, uint16_t *dmin_out) {
    uint8_t Laux[32];
    float weights[32];
    float mins[QK_K / 32];
    float scales[QK_K / 32];
    float max_scale = 0;
    float max_min = 0;
    for (int j = 0; j < QK_K / 32; ++j) {
        float sum_x2 = 0;
        for (int l = 0; l < 32; ++l) sum_x2 += x[32 * j + l] * x[32 * j + l];
        float av_x = sqrtf(sum_x2 / 32);
        for (int l = 0; l < 32; ++l) weights[l] = av_x + fabsf(x[32 * j + l]);
        scales[j] = make_qkx2_quants(32, nmax, x + 32 * j, weights, L + 32 * j, &mins[j], Laux, rmin,
                                     rdelta, nstep, 0);
        float scale = scales[j];
        if (scale > max_scale) max_scale = scale;
        float min = mins[j];
        if (min > max_min) max_min = min;
    }
    float inv_scale = max_scale > 0 ? 63.f / max_scale : 0.f;
    float inv_min = max_min > 0 ? 63.f / max_min : 0.f;
    memset(scales_out, 0, 12);
    for (int j = 0; j < QK_K / 32; ++j) {
        uint8_t ls = (uint8_t)nearest_int(inv_scale * scales[j]);
        uint8_t lm = (uint8_t)nearest_int(inv_min * mins[j]);
        ls = MINI(63, ls);
        lm = MINI(63, lm);
        if (j < 4) {
            scales_out[j] = ls;
            scales_out[j + 4] = lm;
        } else {
            scales_out[j + 4] = (ls & 0xF) | ((lm & 0xF) << 4);
            scales_out[j - 4] |= ((ls >> 4) << 6);
            scales_out[j - 0] |= ((lm >> 4) << 6);
        }
    }
    *d_out = f32_to_f16(max_scale / 63.f);
    *dmin_out = f32_to_f16(max_min / 63.f);
    uint8_t sc, m;
    for (int j = 0; j < QK_K / 32; ++j) {
        get_scale_min_k4(j, scales_out, &sc, &m);
        const float d = f16_to_f32(*d_out) * sc;
        if (!d) continue;
        const float dm = f16_to_f32(*dmin_out) * m;
        for (int ii = 0; ii < 32; ++ii) {
            int l = nearest_int((x[32 * j + ii] + dm) / d);
            l = MAXI(0, MINI(nmax, l));
            L[32 * j + ii] = (uint8_t)l;
        }
    }
}

static void row_q4_K(const float *x, block_q4_K *y, int64_t k) {
    const int64_t nb = k / QK_K;
    uint8_t L[QK_K];
    for (int64_t i = 0; i < nb; i++) {
        k45_scales(x, 15, -1.f, 0.1f, 20, L, y[i].scales, &y[i].d, &y[i].dmin);
        uint8_t *q = y[i].qs;
        for (int j = 0; j < QK_K; j += 64) {
            for (int l = 0; l < 32; ++l) q[l] = L[j + l] | (uint8_t)(L[j + l + 32] << 4);
            q += 32;
        }
        x += QK_K;
    }
}

static void row_q5_K(const float *x, block_q5_K *y, int64_t k) {
    const int64_t nb = k / QK_K;
    uint8_t L[QK_K];
    for (int64_t i = 0; i < nb; i++) {
        k45_scales(x, 31, -0.5f, 0.1f, 15, L, y[i].scales, &y[i].d, &y[i].dmin);
        uint8_t *qh = y[i].qh;
        uint8_t *ql = y[i].qs;
        memset(qh, 0, QK_K / 8);
        uint8_t m1 = 1, m2 = 2;
        for (int n = 0; n < QK_K; n += 64) {
            for (int j = 0; j < 32; ++j) {
                int l1 = L[n + j];
                if (l1 > 15) { l1 -= 16; qh[j] |= m1; }
                int l2 = L[n + j + 32];
                if (l2 > 15) { l2 -= 16; qh[j] |= m2; }
                ql[j] = (uint8_t)(l1 | (l2 << 4));
            }
            m1 <<= 2;
            m2 <<= 2;
            ql += 32;
        }
        x += QK_K;
    }
}

/* ---- Q2_K: 16 sub-blocks x 16, 2-bit codes, 4-bit scale + 4-bit min per sub-block -------------
 * llama.cpp quantize_row_q2_K_ref: make_qkx2_quants(16, 3, x, |x|, .., -0.5, 0.1, 15, use_mad=true). */
static void row_q2_K(const float *x, block_q2_K *y, int64_t k) {
    const int64_t nb = k / QK_K;
    uint8_t L[QK_K];
    uint8_t Laux[16];
    float weights[16];
    float mins[QK_K / 16];
    float scales[QK_K / 16];
    const float q4scale = 15.f;
    for (int64_t i = 0; i < nb; i++) {
        float max_scale = 0;
        float max_min = 0;
        for (int j = 0; j < QK_K / 16; ++j) {
            for (int l = 0; l < 16; ++l) weights[l] = fabsf(x[16 * j + l]);
            scales[j] = make_qkx2_quants(16, 3, x + 16 * j, weights, L + 16 * j, &mins[j], Laux, -0.5f, 0.1f, 15, 1);
            float scale = scales[j];
            if (scale > max_scale) max_scale = scale;
            float min = mins[j];
            if (min > max_min) max_min = min;
        }
        if (max_scale > 0) {
            float iscale = q4scale / max_scale;
            for (int j = 0; j < QK_K / 16; ++j) {
                int l = nearest_int(iscale * scales[j]);
                y[i].scales[j] = (uint8_t)l;
            }
            y[i].d = f32_to_f16(max_scale / q4scale);
        } else {
            for (int j = 0; j < QK_K / 16; ++j) y[i].scales[j] = 0;
            y[i].d = f32_to_f16(0.f);
        }
        if (max_min > 0) {
            float iscale = q4scale / max_min;
            for (int j = 0; j < QK_K / 16; ++j) {
                int l = nearest_int(iscale * mins[j]);
                y[i].scales[j] |= (uint8_t)(l << 4);
            }
            y[i].dmin = f32_to_f16(max_min / q4scale);
        } else {
            y[i].dmin = f32_to_f16(0.f);
        }
        for (int j = 0; j < QK_K / 16; ++j) {
            const float d = f16_to_f32(y[i].d) * (y[i].scales[j] & 0xF);
            if (!d) continue;
            const float dm = f16_to_f32(y[i].dmin) * (y[i].scales[j] >> 4);
            for (int ii = 0; ii < 16; ++ii) {
                int l = nearest_int((x[16 * j + ii] + dm) / d);
                l = MAXI(0, MINI(3, l));
                L[16 * j + ii] = (uint8_t)l;
            }
        }
        for (int j = 0; j < QK_K; j += 128) {
            for (int l = 0; l < 32; ++l) {
                y[i].qs[j / 4 + l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 2) | (L[j + l + 64] << 4) | (L[j + l + 96] << 6));
            }
        }
        x += QK_K;
    }
}

/* ---- Q3_K: 16 sub-blocks x 16, 3-bit codes (2 low bits + high-bit mask), 6-bit signed scales ---
 * llama.cpp make_q3_quants(n=16, nmax=4, do_rmse=true): greedy start at -nmax/max, then up to five
 * sweeps of coordinate descent on the x^2-weighted fit. */
static float make_q3_quants(int n, int nmax, const float *x, int8_t *L, int do_rmse) {
    float max = 0;
    float amax = 0;
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; max = x[i]; }
    }
    if (amax < GROUP_MAX_EPS) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        return 0.f;
    }
    float iscale = -nmax / max;
    if (do_rmse) {
        float sumlx = 0;
        float suml2 = 0;
        for (int i = 0; i < n; ++i) {
            int l = nearest_int(iscale * x[i]);
            l = MAXI(-nmax, MINI(nmax - 1, l));
            L[i] = (int8_t)l;
            float w = x[i] * x[i];
            sumlx += w * x[i] * l;
            suml2 += w * l * l;
        }
        for (int itry = 0; itry < 5; ++itry) {
            int n_changed = 0;
            for (int i = 0; i < n; ++i) {
                float w = x[i] * x[i];
                float slx = sumlx - w * x[i] * L[i];
                if (slx > 0) {
                    float sl2 = suml2 - w * L[i] * L[i];
                    int new_l = nearest_int(x[i] * sl2 / slx);
                    new_l = MAXI(-nmax, MINI(nmax - 1, new_l));
                    if (new_l != L[i]) {
                        slx += w * x[i] * new_l;
                        sl2 += w * new_l * new_l;
                        if (sl2 > 0 && slx * slx * suml2 > sumlx * sumlx * sl2) {
                            L[i] = (int8_t)new_l; sumlx = slx; suml2 = sl2;
                            ++n_changed;
                        }
                    }
                }
            }
            if (!n_changed) break;
        }
        for (int i = 0; i < n; ++i) L[i] += nmax;
        return sumlx / suml2;
    }
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * x[i]);
        l = MAXI(-nmax, MINI(nmax - 1, l));
        L[i] = (int8_t)(l + nmax);
    }
    return 1 / iscale;
}

static void row_q3_K(const float *x, block_q3_K *y, int64_t k) {
    const int64_t nb = k / QK_K;
    int8_t L[QK_K];
    float scales[QK_K / 16];
    for (int64_t i = 0; i < nb; i++) {
        float max_scale = 0;
        float amax = 0;
        for (int j = 0; j < QK_K / 16; ++j) {
            scales[j] = make_q3_quants(16, 4, x + 16 * j, L + 16 * j, 1);
            float scale = fabsf(scales[j]);
            if (scale > amax) { amax = scale; max_scale = scales[j]; }
        }
        memset(y[i].scales, 0, 12);
        if (max_scale) {
            float iscale = -32.f / max_scale;
            for (int j = 0; j < QK_K / 16; ++j) {
                int8_t l = (int8_t)nearest_int(iscale * scales[j]);
                l = (int8_t)(MAXI(-32, MINI(31, l)) + 32);
                if (j < 8) {
                    y[i].scales[j] = l & 0xF;
                } else {
                    y[i].scales[j - 8] |= ((l & 0xF) << 4);
                }
                l >>= 4;
                y[i].scales[j % 4 + 8] |= (l << (2 * (j / 4)));
            }
            y[i].d = f32_to_f16(1 / iscale);
        } else {
            y[i].d = f32_to_f16(0.f);
        }
        int8_t sc;
        for (int j = 0; j < QK_K / 16; ++j) {
            sc = j < 8 ? y[i].scales[j] & 0xF : y[i].scales[j - 8] >> 4;
            sc = (int8_t)((sc | (((y[i].scales[8 + j % 4] >> (2 * (j / 4))) & 3) << 4)) - 32);
            float d = f16_to_f32(y[i].d) * sc;
            if (!d) continue;
            for (int ii = 0; ii < 16; ++ii) {
                int l = nearest_int(x[16 * j + ii] / d);
                l = MAXI(-4, MINI(3, l));
                L[16 * j + ii] = (int8_t)(l + 4);
            }
        }
        memset(y[i].hmask, 0, QK_K / 8);
        /* high bit of the first 32 codes -> bit 0 of hmask[0..31], the next 32 -> bit 1, ... */
        int m = 0;
        uint8_t hm = 1;
        for (int j = 0; j < QK_K; ++j) {
            if (L[j] > 3) {
                y[i].hmask[m] |= hm;
                L[j] -= 4;
            }
            if (++m == QK_K / 8) { m = 0; hm <<= 1; }
        }
        for (int j = 0; j < QK_K; j += 128) {
            for (int l = 0; l < 32; ++l) {
                y[i].qs[j / 4 + l] = (uint8_t)(L[j + l] | (L[j + l + 32] << 2) | (L[j + l + 64] << 4) | (L[j + l + 96] << 6));
            }
        }
        x += QK_K;
    }
}

/* ---- IQ4_NL: 32-element blocks, 4-bit indices into a fixed non-linear table ------------------------
 * llama.cpp falls back to it for Q2_K / Q3_K tensors whose rows are not a multiple of 256.  llama-quantize
 * reaches it through quantize_iq4_nl -> quantize_row_iq4_nl_impl(super_block = block = 32, quant_weights = NULL,
 * ntry = 7): weights x^2, first guess d = -max/values[0], then 15 candidate inverse scales (itry + values[0])/max,
 * keep the one maximising (sum w q x)^2 / (sum w q^2); indices are re-derived from the final fp32 scale. */
static const int8_t kvalues_iq4nl[16] = {-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113};

static inline int best_index_int8(int n, const int8_t *val, float x) {
    if (x <= val[0]) return 0;
    if (x >= val[n - 1]) return n - 1;
    int ml = 0, mu = n - 1;
    while (mu - ml > 1) {
        int mav = (ml + mu) / 2;
        if (x < val[mav]) mu = mav; else ml = mav;
    }
    return x - val[mu - 1] < val[mu] - x ? mu - 1 : mu;
}

static void row_iq4_nl(const float *x, block_iq4_nl *y, int64_t k) {
    const int64_t nb = k / QK;
    const int8_t *values = kvalues_iq4nl;
    const int ntry = 7;
    for (int64_t ib = 0; ib < nb; ib++) {
        const float *xb = x + QK * ib;
        uint8_t L[QK];
        float weight[QK];
        memset(y[ib].qs, 0, QK / 2);
        y[ib].d = f32_to_f16(0.f);
        for (int j = 0; j < QK; ++j) weight[j] = xb[j] * xb[j];
        float amax = 0, max = 0;
        for (int j = 0; j < QK; ++j) {
            float ax = fabsf(xb[j]);
            if (ax > amax) { amax = ax; max = xb[j]; }
        }
        float scale = 0;
        if (!(amax < GROUP_MAX_EPS)) {
            float d = ntry > 0 ? -max / values[0] : max / values[0];
            float id = 1 / d;
            float sumqx = 0, sumq2 = 0;
            for (int j = 0; j < QK; ++j) {
                float al = id * xb[j];
                int l = best_index_int8(16, values, al);
                float q = values[l];
                float w = weight[j];
                sumqx += w * q * xb[j];
                sumq2 += w * q * q;
            }
            d = sumqx / sumq2;
            float best = d * sumqx;
            for (int itry = -ntry; itry <= ntry; ++itry) {
                id = (itry + values[0]) / max;
                sumqx = sumq2 = 0;
                for (int j = 0; j < QK; ++j) {
                    float al = id * xb[j];
                    int l = best_index_int8(16, values, al);
                    float q = values[l];
                    float w = weight[j];
                    sumqx += w * q * xb[j];
                    sumq2 += w * q * q;
                }
                if (sumq2 > 0 && sumqx * sumqx > best * sumq2) {
                    d = sumqx / sumq2; best = d * sumqx;
                }
            }
            scale = d;
        }
        y[ib].d = f32_to_f16(scale);
        float id = scale ? 1 / scale : 0;
        for (int j = 0; j < QK; ++j) L[j] = (uint8_t)best_index_int8(16, values, id * xb[j]);
        for (int j = 0; j < 16; ++j) y[ib].qs[j] = (uint8_t)(L[j] | (L[16 + j] << 4));
    }
}

/* ---- D.5 Q6_K ------------------------------------------------------------------- */
static float make_qx_quants_rmse1(int n, int nmax, const float *x, int8_t *L) {
    float max = 0;
    float amax = 0;
    for (int i = 0; i < n; ++i) {
        float ax = fabsf(x[i]);
        if (ax > amax) { amax = ax; max = x[i]; }
    }
    if (amax < GROUP_MAX_EPS) {
        for (int i = 0; i < n; ++i) L[i] = 0;
        return 0.f;
    }
    float iscale = -nmax / max;
    float sumlx = 0;
    float suml2 = 0;
    for (int i = 0; i < n; ++i) {
        int l = nearest_int(iscale * x[i]);
        l = MAXI(-nmax, MINI(nmax - 1, l));
        L[i] = (int8_t)(l + nmax);
        float w = x[i] * x[i];
        sumlx += w * x[i] * l;
        suml2 += w * l * l;
    }
    float scale = suml2 ? sumlx / suml2 : 0.0f;
    float best = scale * sumlx;
    for (int is = -9; is <= 9; ++is) {
        if (is == 0) continue;
        iscale = -(nmax + 0.1f * is) / max;
        sumlx = suml2 = 0;
        for (int i = 0; i < n; ++i) {
            int l = nearest_int(iscale * x[i]);
            l = MAXI(-nmax, MINI(nmax - 1, l));
            float w = x[i] * x[i];
            sumlx += w * x[i] * l;
            suml2 += w * l * l;
        }
        if (suml2 > 0 && sumlx * sumlx > best * suml2) {
            for (int i = 0; i < n; ++i) {
                int l = nearest_int(iscale * x[i]);
                L[i] = (int8_t)(nmax + MAXI(-nmax, MINI(nmax - 1, l)));
            }
            scale = sumlx / suml2;
            best = scale * sumlx;
        }
    }
    return scale;
}

static void row_q6_K(const float *x, block_q6_K *y, int64_t k) {
    const int64_t nb = k / QK_K;
    int8_t L[QK_K];
    float scales[QK_K / 16];
    for (int64_t i = 0; i < nb; i++) {
        float max_scale = 0;
        float max_abs_scale = 0;
        for (int ib = 0; ib < QK_K / 16; ++ib) {
            const float scale = make_qx_quants_rmse1(16, 32, x + 16 * ib, L + 16 * ib);
            scales[ib] = scale;
            const float abs_scale = fabsf(scale);
            if (abs_scale > max_abs_scale) {
                max_abs_scale = abs_scale;
                max_scale = scale;
            }
        }
        if (max_abs_scale < GROUP_MAX_EPS) {
            memset(&y[i], 0, sizeof(block_q6_K));
            y[i].d = f32_to_f16(0.f);
            x += QK_K;
            continue;
        }
        float iscale = -128.f / max_scale;
        y[i].d = f32_to_f16(1 / iscale);
        for (int ib = 0; ib < QK_K / 16; ++ib)
            y[i].scales[ib] = (int8_t)MINI(127, nearest_int(iscale * scales[ib]));
        for (int j = 0; j < QK_K / 16; ++j) {
            float d = f16_to_f32(y[i].d) * y[i].scales[j];
            if (!d) continue;
            for (int ii = 0; ii < 16; ++ii) {
                int l = nearest_int(x[16 * j + ii] / d);
                l = MAXI(-32, MINI(31, l));
                L[16 * j + ii] = (int8_t)(l + 32);
            }
        }
        uint8_t *ql = y[i].ql;
        uint8_t *qh = y[i].qh;
        for (int j = 0; j < QK_K; j += 128) {
            for (int l = 0; l < 32; ++l) {
                const uint8_t q1 = L[j + l + 0] & 0xF;
                const uint8_t q2 = L[j + l + 32] & 0xF;
                const uint8_t q3 = L[j + l + 64] & 0xF;
                const uint8_t q4 = L[j + l + 96] & 0xF;
                ql[l + 0] = q1 | (uint8_t)(q3 << 4);
                ql[l + 32] = q2 | (uint8_t)(q4 << 4);
                qh[l] = (uint8_t)((L[j + l] >> 4) | ((L[j + l + 32] >> 4) << 2) |
                                  ((L[j + l + 64] >> 4) << 4) | ((L[j + l + 96] >> 4) << 6));
            }
            ql += 64;
            qh += 32;
        }
        x += QK_K;
    }
}

/* ---- dequantizers (llama.cpp dequantize_row_*; twins GGUFPY/quants.py:241-572) ----- */
static void deq_q8_0(const block_q8_0 *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float d = f16_to_f32(x[i].d);
        for (int j = 0; j < QK; ++j) y[i * QK + j] = x[i].qs[j] * d;
    }
}
static void deq_q4_0(const block_q4_0 *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float d = f16_to_f32(x[i].d);
        for (int j = 0; j < QK / 2; ++j) {
            const int x0 = (x[i].qs[j] & 0x0F) - 8;
            const int x1 = (x[i].qs[j] >> 4) - 8;
            y[i * QK + j + 0] = x0 * d;
            y[i * QK + j + QK / 2] = x1 * d;
        }
    }
}
static void deq_q4_1(const block_q4_1 *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float d = f16_to_f32(x[i].d), m = f16_to_f32(x[i].m);
        for (int j = 0; j < QK / 2; ++j) {
            y[i * QK + j + 0] = (x[i].qs[j] & 0x0F) * d + m;
            y[i * QK + j + QK / 2] = (x[i].qs[j] >> 4) * d + m;
        }
    }
}
static void deq_q5_0(const block_q5_0 *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float d = f16_to_f32(x[i].d);
        uint32_t qh;
        memcpy(&qh, x[i].qh, 4);
        for (int j = 0; j < QK / 2; ++j) {
            const uint8_t xh_0 = ((qh >> (j + 0)) << 4) & 0x10;
            const uint8_t xh_1 = ((qh >> (j + 12))) & 0x10;
            const int32_t x0 = ((x[i].qs[j] & 0x0F) | xh_0) - 16;
            const int32_t x1 = ((x[i].qs[j] >> 4) | xh_1) - 16;
            y[i * QK + j + 0] = x0 * d;
            y[i * QK + j + QK / 2] = x1 * d;
        }
    }
}
static void deq_q5_1(const block_q5_1 *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK; i++) {
        const float d = f16_to_f32(x[i].d), m = f16_to_f32(x[i].m);
        uint32_t qh;
        memcpy(&qh, x[i].qh, 4);
        for (int j = 0; j < QK / 2; ++j) {
            const uint8_t xh_0 = ((qh >> (j + 0)) << 4) & 0x10;
            const uint8_t xh_1 = ((qh >> (j + 12))) & 0x10;
            const int x0 = (x[i].qs[j] & 0x0F) | xh_0;
            const int x1 = (x[i].qs[j] >> 4) | xh_1;
            y[i * QK + j + 0] = x0 * d + m;
            y[i * QK + j + QK / 2] = x1 * d + m;
        }
    }
}
static void deq_iq4_nl(const block_iq4_nl *x, float *y, int64_t k) {
    const int64_t nb = k / QK;
    for (int64_t i = 0; i < nb; i++) {
        const float d = f16_to_f32(x[i].d);
        for (int j = 0; j < 16; ++j) {
            y[i * QK + j] = d * kvalues_iq4nl[x[i].qs[j] & 0xF];
            y[i * QK + j + 16] = d * kvalues_iq4nl[x[i].qs[j] >> 4];
        }
    }
}

static void deq_q2_K(const block_q2_K *x, float *y, int64_t k) {
    const int64_t nb = k / QK_K;
    for (int64_t i = 0; i < nb; i++) {
        const float d = f16_to_f32(x[i].d);
        const float min = f16_to_f32(x[i].dmin);
        const uint8_t *q = x[i].qs;
        int is = 0;
        for (int n = 0; n < QK_K; n += 128) {
            int shift = 0;
            for (int j = 0; j < 4; ++j) {
                uint8_t sc = x[i].scales[is++];
                float dl = d * (sc & 0xF), ml = min * (sc >> 4);
                for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l] >> shift) & 3)) - ml;
                sc = x[i].scales[is++];
                dl = d * (sc & 0xF); ml = min * (sc >> 4);
                for (int l = 0; l < 16; ++l) *y++ = dl * ((int8_t)((q[l + 16] >> shift) & 3)) - ml;
                shift += 2;
            }
            q += 32;
        }
    }
}

static void deq_q3_K(const block_q3_K *x, float *y, int64_t k) {
    const int64_t nb = k / QK_K;
    for (int64_t i = 0; i < nb; i++) {
        const float d_all = f16_to_f32(x[i].d);
        const uint8_t *q = x[i].qs;
        const uint8_t *hm = x[i].hmask;
        uint8_t m = 1;
        int8_t scales[16];
        for (int j = 0; j < 16; ++j) {
            int lo = j < 8 ? x[i].scales[j] & 0xF : x[i].scales[j - 8] >> 4;
            int hi = (x[i].scales[8 + j % 4] >> (2 * (j / 4))) & 3;
            scales[j] = (int8_t)((lo | (hi << 4)) - 32);
        }
        int is = 0;
        for (int n = 0; n < QK_K; n += 128) {
            int shift = 0;
            for (int j = 0; j < 4; ++j) {
                float dl = d_all * scales[is++];
                for (int l = 0; l < 16; ++l)
                    *y++ = dl * ((int8_t)((q[l + 0] >> shift) & 3) - ((hm[l + 0] & m) ? 0 : 4));
                dl = d_all * scales[is++];
                for (int l = 0; l < 16; ++l)
                    *y++ = dl * ((int8_t)((q[l + 16] >> shift) & 3) - ((hm[l + 16] & m) ? 0 : 4));
                shift += 2;
                m <<= 1;
            }
            q += 32;
        }
    }
}

static void deq_q4_K(const block_q4_K *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK_K; i++) {
        const uint8_t *q = x[i].qs;
        const float d = f16_to_f32(x[i].d), min = f16_to_f32(x[i].dmin);
        int is = 0;
        uint8_t sc, m;
        for (int j = 0; j < QK_K; j += 64) {
            get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
            const float d1 = d * sc, m1 = min * m;
            get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
            const float d2 = d * sc, m2 = min * m;
            for (int l = 0; l < 32; ++l) *y++ = d1 * (q[l] & 0xF) - m1;
            for (int l = 0; l < 32; ++l) *y++ = d2 * (q[l] >> 4) - m2;
            q += 32;
            is += 2;
        }
    }
}
static void deq_q5_K(const block_q5_K *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK_K; i++) {
        const uint8_t *ql = x[i].qs;
        const uint8_t *qh = x[i].qh;
        const float d = f16_to_f32(x[i].d), min = f16_to_f32(x[i].dmin);
        int is = 0;
        uint8_t sc, m;
        uint8_t u1 = 1, u2 = 2;
        for (int j = 0; j < QK_K; j += 64) {
            get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
            const float d1 = d * sc, m1 = min * m;
            get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
            const float d2 = d * sc, m2 = min * m;
            for (int l = 0; l < 32; ++l) *y++ = d1 * ((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)) - m1;
            for (int l = 0; l < 32; ++l) *y++ = d2 * ((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)) - m2;
            ql += 32;
            is += 2;
            u1 <<= 2;
            u2 <<= 2;
        }
    }
}
static void deq_q6_K(const block_q6_K *x, float *y, int64_t k) {
    for (int64_t i = 0; i < k / QK_K; i++) {
        const float d = f16_to_f32(x[i].d);
        const uint8_t *ql = x[i].ql;
        const uint8_t *qh = x[i].qh;
        const int8_t *sc = x[i].scales;
        for (int n = 0; n < QK_K; n += 128) {
            for (int l = 0; l < 32; ++l) {
                int is = l / 16;
                const int8_t q1 = (int8_t)((ql[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                const int8_t q3 = (int8_t)((ql[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                y[l + 0] = d * sc[is + 0] * q1;
                y[l + 32] = d * sc[is + 2] * q2;
                y[l + 64] = d * sc[is + 4] * q3;
                y[l + 96] = d * sc[is + 6] * q4;
            }
            y += 128;
            ql += 64;
            qh += 32;
            sc += 8;
        }
    }
}

/* ---- exported entry points --------------------------------------------------------- */
/* ggml_type ids (GGUFPY/constants.py GGMLQuantizationType) */
enum { T_IQ4_NL = 20, T_Q2_K = 10, T_Q3_K = 11, T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14 };

int oracle_block_elems(int t) {
    switch (t) {
        case T_Q4_0: case T_Q4_1: case T_Q5_0: case T_Q5_1: case T_Q8_0: case T_IQ4_NL: return QK;
        case T_Q2_K: case T_Q3_K: case T_Q4_K: case T_Q5_K: case T_Q6_K: return QK_K;
        default: return -1;
    }
}
int oracle_block_bytes(int t) {
    switch (t) {
        case T_Q4_0: case T_IQ4_NL: return 18; case T_Q4_1: return 20; case T_Q5_0: return 22; case T_Q5_1: return 24;
        case T_Q8_0: return 34; case T_Q2_K: return 84; case T_Q3_K: return 110; case T_Q4_K: return 144; case T_Q5_K: return 176; case T_Q6_K: return 210;
        default: return -1;
    }
}

/* ---- pthread parallel-for over rows (libgomp is not in this image) ------------------ */
#include <pthread.h>
#include <unistd.h>

static int g_threads = 0; /* 0 = all online cores */
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef void (*range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int64_t lo, hi; } job_t;
static void *job_main(void *p) {
    job_t *j = (job_t *)p;
    j->fn(j->lo, j->hi, j->ctx);
    return 0;
}
static void parallel_for(int64_t n, range_fn fn, void *ctx) {
    int nt = oracle_get_threads();
    if (nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t th[256];
    job_t jobs[256];
    if (nt > 256) nt = 256;
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / nt; jobs[t].hi = n * (t + 1) / nt;
        pthread_create(&th[t], 0, job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], 0);
}

typedef struct { int t; const void *src; void *dst; int64_t ncols, row_bytes; } qctx_t;

/* x: [nrows, ncols] fp32 row-major; y: [nrows, ncols/elems*bytes].  Rows are split over
 * threads the way llama-quantize splits chunks over nthread workers:
 * rows/blocks are independent so the bytes do not depend on the thread count. */
static void quant_rows(int64_t lo, int64_t hi, void *p) {
    qctx_t *c = (qctx_t *)p;
    for (int64_t r = lo; r < hi; r++) {
        const float *xr = (const float *)c->src + r * c->ncols;
        uint8_t *yr = (uint8_t *)c->dst + r * c->row_bytes;
        switch (c->t) {
            case T_Q4_0: row_q4_0(xr, (block_q4_0 *)yr, c->ncols); break;
            case T_Q4_1: row_q4_1(xr, (block_q4_1 *)yr, c->ncols); break;
            case T_Q5_0: row_q5_0(xr, (block_q5_0 *)yr, c->ncols); break;
            case T_Q5_1: row_q5_1(xr, (block_q5_1 *)yr, c->ncols); break;
            case T_Q8_0: row_q8_0(xr, (block_q8_0 *)yr, c->ncols); break;
            case T_IQ4_NL: row_iq4_nl(xr, (block_iq4_nl *)yr, c->ncols); break;
            case T_Q2_K: row_q2_K(xr, (block_q2_K *)yr, c->ncols); break;
            case T_Q3_K: row_q3_K(xr, (block_q3_K *)yr, c->ncols); break;
            case T_Q4_K: row_q4_K(xr, (block_q4_K *)yr, c->ncols); break;
            case T_Q5_K: row_q5_K(xr, (block_q5_K *)yr, c->ncols); break;
            case T_Q6_K: row_q6_K(xr, (block_q6_K *)yr, c->ncols); break;
        }
    }
}
static void deq_rows(int64_t lo, int64_t hi, void *p) {
    qctx_t *c = (qctx_t *)p;
    for (int64_t r = lo; r < hi; r++) {
        const uint8_t *xr = (const uint8_t *)c->src + r * c->row_bytes;
        float *yr = (float *)c->dst + r * c->ncols;
        switch (c->t) {
            case T_Q4_0: deq_q4_0((const block_q4_0 *)xr, yr, c->ncols); break;
            case T_Q4_1: deq_q4_1((const block_q4_1 *)xr, yr, c->ncols); break;
            case T_Q5_0: deq_q5_0((const block_q5_0 *)xr, yr, c->ncols); break;
            case T_Q5_1: deq_q5_1((const block_q5_1 *)xr, yr, c->ncols); break;
            case T_Q8_0: deq_q8_0((const block_q8_0 *)xr, yr, c->ncols); break;
            case T_IQ4_NL: deq_iq4_nl((const block_iq4_nl *)xr, yr, c->ncols); break;
            case T_Q2_K: deq_q2_K((const block_q2_K *)xr, yr, c->ncols); break;
            case T_Q3_K: deq_q3_K((const block_q3_K *)xr, yr, c->ncols); break;
            case T_Q4_K: deq_q4_K((const block_q4_K *)xr, yr, c->ncols); break;
            case T_Q5_K: deq_q5_K((const block_q5_K *)xr, yr, c->ncols); break;
            case T_Q6_K: deq_q6_K((const block_q6_K *)xr, yr, c->ncols); break;
        }
    }
}

int oracle_quantize(int t, const float *x, void *y, int64_t nrows, int64_t ncols) {
    const int be = oracle_block_elems(t), bb = oracle_block_bytes(t);
    if (be < 0 || ncols % be) return -1;
    qctx_t c = {t, x, y, ncols, ncols / be * bb};
    parallel_for(nrows, quant_rows, &c);
    return 0;
}

int oracle_dequantize(int t, const void *x, float *y, int64_t nrows, int64_t ncols) {
    const int be = oracle_block_elems(t), bb = oracle_block_bytes(t);
    if (be < 0 || ncols % be) return -1;
    qctx_t c = {t, x, y, ncols, ncols / be * bb};
    parallel_for(nrows, deq_rows, &c);
    return 0;
}

/* fp16 round trip of an fp32 buffer: the reference always goes HF -> f16 GGUF -> quantize
 * (SURVEY §3.2 note: "weights are rounded to fp16 first") */
typedef struct { const void *src; float *dst; } cctx_t;
static void round_rng(int64_t lo, int64_t hi, void *p) {
    cctx_t *c = (cctx_t *)p;
    for (int64_t i = lo; i < hi; i++) c->dst[i] = f16_to_f32(f32_to_f16(((const float *)c->src)[i]));
}
static void cvt_rng(int64_t lo, int64_t hi, void *p) {
    cctx_t *c = (cctx_t *)p;
    for (int64_t i = lo; i < hi; i++) c->dst[i] = f16_to_f32(((const uint16_t *)c->src)[i]);
}
void oracle_round_f16(const float *x, float *y, int64_t n) {
    cctx_t c = {x, y};
    parallel_for(n, round_rng, &c);
}
void oracle_f16_to_f32(const uint16_t *x, float *y, int64_t n) {
    cctx_t c = {x, y};
    parallel_for(n, cvt_rng, &c);
}
