// AWQ scale-search kernels on sm_100a.
//
// Replaces UPSTREAM llmcompressor AWQModifier `_compute_best_scale`, `_pseudo_quantize_tensor`,
// `_compute_loss` reached from
// ref/src/quantool/methods/llm_compressor/awq/awq.py:81 via llmcompressor.oneshot at
// ref/src/quantool/methods/llm_compressor/base.py:159-161.
//
//   awq_group_absmax / awq_wmean : w_mean = column mean of |W| normalised by its group's max
//   awq_scale_qdq                : W' = pseudo_quant(W * s) / s in ONE pass over W per grid point
//                                  (scale -> group min/max -> quantize -> dequantize -> unscale)
//   sq_err_sum                   : sum((a - b)^2) of the parent outputs, fp32 squares, fp64 total
//
// torch evaluates these on the model-dtype weight: every elementwise op is "fp32 compute,
// round to the tensor dtype".  DT reproduces that rounding after each op.  All HBM-bound.
#include "common.cuh"

namespace qt {
namespace awq {

template <int DT>
QT_D float rnd(float v) {
    if (DT == QT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (DT == QT_F16) return __half2float(__float2half_rn(v));
    return v;
}
template <int DT>
QT_D float ld(const void* p, long long i) {
    if (DT == QT_F32) return reinterpret_cast<const float*>(p)[i];
    if (DT == QT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <int DT>
QT_D void st(void* p, long long i, float v) {
    if (DT == QT_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (DT == QT_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// one warp per row, loop over groups: gamax[n][g] = max |W[n][g*gs .. (g+1)*gs)|
template <int DT>
__global__ void __launch_bounds__(256) group_absmax_kernel(const void* __restrict__ W, int N, int K, int gs,
                                                           float* __restrict__ gamax) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    const int G = K / gs;
    for (int g = 0; g < G; g++) {
        float m = 0.f;
        for (int c = lane; c < gs; c += 32) m = fmaxf(m, fabsf(ld<DT>(W, (long long)row * K + (long long)g * gs + c)));
        m = warp_max(m);
        if (lane == 0) gamax[(long long)row * G + g] = m;
    }
}

// colsum[c] += sum_n rnd(|W[n][c]| / rnd(gamax[n][c/gs] + 1e-6))
template <int DT>
__global__ void __launch_bounds__(256) wmean_kernel(const void* __restrict__ W, const float* __restrict__ gamax,
                                                    int N, int K, int gs, float* __restrict__ colsum) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= K) return;
    const int G = K / gs, g = c / gs;
    float acc = 0.f;
    for (int n = blockIdx.y; n < N; n += gridDim.y) {
        const float den = rnd<DT>(gamax[(long long)n * G + g] + rnd<DT>(1e-6f));
        acc += rnd<DT>(fabsf(ld<DT>(W, (long long)n * K + c)) / den);
    }
    atomicAdd(colsum + c, acc);
}

// 4 consecutive elements of a 16/32-bit row
template <int DT>
QT_D void ld4(const void* p, long long i, float v[4]) {
    if (DT == QT_F32) {
        const float4 a = *reinterpret_cast<const float4*>((const float*)p + i);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
        const uint2 a = *reinterpret_cast<const uint2*>((const char*)p + i * 2);
        if (DT == QT_F16) {
            v[0] = f16_bits_to_float(a.x & 0xffffu); v[1] = f16_bits_to_float(a.x >> 16);
            v[2] = f16_bits_to_float(a.y & 0xffffu); v[3] = f16_bits_to_float(a.y >> 16);
        } else {
            v[0] = bf16_bits_to_float(a.x & 0xffffu); v[1] = bf16_bits_to_float(a.x >> 16);
            v[2] = bf16_bits_to_float(a.y & 0xffffu); v[3] = bf16_bits_to_float(a.y >> 16);
        }
    }
}
template <int DT>
QT_D void st4(void* p, long long i, const float v[4]) {
    if (DT == QT_F32) {
        *reinterpret_cast<float4*>((float*)p + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        uint2 o;
        if (DT == QT_F16) {
            o.x = (uint32_t)__half_as_ushort(__float2half_rn(v[0])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[1])) << 16);
            o.y = (uint32_t)__half_as_ushort(__float2half_rn(v[2])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[3])) << 16);
        } else {
            o.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[0])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[1])) << 16);
            o.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[3])) << 16);
        }
        *reinterpret_cast<uint2*>((char*)p + i * 2) = o;
    }
}

// out[n][c] = rnd( pq( rnd(W[n][c] * s[c]) ) / s[c] )   (pq = _pseudo_quantize_tensor per group)
// One warp per row; a group of gs columns is walked in chunks of 128 (lane = 4 consecutive
// columns, 64/128-bit accesses); the scaled values stay in registers when gs <= 128, which is the
// preset case (one pass over W), and are recomputed from L1 for larger groups.
template <int DT>
__global__ void __launch_bounds__(256) scale_qdq_kernel(const void* __restrict__ W, const float* __restrict__ s,
                                                        int N, int K, int gs, int num_bits, int symmetric,
                                                        void* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    const int G = K / gs;
    const float max_int_s = (float)((1 << (num_bits - 1)) - 1), min_int_s = -(float)(1 << (num_bits - 1));
    const float max_int_a = (float)((1 << num_bits) - 1);
    for (int g = 0; g < G; g++) {
        const long long base = (long long)row * K + (long long)g * gs;
        const float* sg = s + (long long)g * gs;
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
        float keep[4], ks[4];
        for (int c0 = lane * 4; c0 < gs; c0 += 128) {
            float w4[4];
            ld4<DT>(W, base + c0, w4);
            const float4 s4 = *reinterpret_cast<const float4*>(sg + c0);
            ks[0] = s4.x; ks[1] = s4.y; ks[2] = s4.z; ks[3] = s4.w;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                keep[i] = rnd<DT>(w4[i] * ks[i]);
                mn = fminf(mn, keep[i]);
                mx = fmaxf(mx, keep[i]);
            }
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        float sc, z = 0.f;
        if (symmetric) {
            const float max_val = fmaxf(fmaxf(fabsf(mn), fabsf(mx)), rnd<DT>(1e-5f));
            sc = rnd<DT>(max_val / max_int_s);
        } else {
            sc = rnd<DT>(fmaxf(rnd<DT>(mx - mn), rnd<DT>(1e-5f)) / max_int_a);
            z = -rintf(rnd<DT>(mn / sc));
            z = fminf(fmaxf(z, 0.f), max_int_a);
        }
        for (int c0 = lane * 4; c0 < gs; c0 += 128) {
            if (gs > 128) {   // values not kept: recompute
                float w4[4];
                ld4<DT>(W, base + c0, w4);
                const float4 s4 = *reinterpret_cast<const float4*>(sg + c0);
                ks[0] = s4.x; ks[1] = s4.y; ks[2] = s4.z; ks[3] = s4.w;
#pragma unroll
                for (int i = 0; i < 4; i++) keep[i] = rnd<DT>(w4[i] * ks[i]);
            }
            float o4[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float dq;
                if (symmetric) {
                    float q = rintf(rnd<DT>(keep[i] / sc));
                    q = fminf(fmaxf(q, min_int_s), max_int_s);
                    dq = rnd<DT>(q * sc);
                } else {
                    float q = rnd<DT>(rintf(rnd<DT>(keep[i] / sc)) + z);
                    q = fminf(fmaxf(q, 0.f), max_int_a);
                    dq = rnd<DT>(rnd<DT>(q - z) * sc);
                }
                o4[i] = dq / ks[i];
            }
            st4<DT>(out, base + c0, o4);
        }
    }
}

// Same arithmetic, group_size 32 / 64 / 128: LPG = group_size / 16 consecutive lanes share a group and every lane owns
// 16 consecutive columns (two 128-bit loads of a 16-bit row), so a warp works on 32 / LPG groups at once with 16
// independent elements per lane - the one-group-per-warp form above is bound by the latency of its serial
// min/max -> divide -> divide chain (12 % of HBM peak, ncu issue 78 %).
// DELTA: instead of the candidate weight, write D = candidate - W (exact in fp32) as bf16 (the MMA operand of the
// Gram-form loss) and as fp32 (its epilogue operand).
template <int DT>
QT_D void ld16(const void* p, long long i, float v[16]) {
    if (DT == QT_F32) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 a = *reinterpret_cast<const float4*>((const float*)p + i + 4 * q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint4 a = *reinterpret_cast<const uint4*>((const char*)p + (i + 8 * q) * 2);
            const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (DT == QT_F16) {
                    v[8 * q + 2 * t] = f16_bits_to_float(u[t] & 0xffffu);
                    v[8 * q + 2 * t + 1] = f16_bits_to_float(u[t] >> 16);
                } else {
                    v[8 * q + 2 * t] = __uint_as_float(u[t] << 16);
                    v[8 * q + 2 * t + 1] = __uint_as_float(u[t] & 0xffff0000u);
                }
            }
        }
    }
}
template <int DT>
QT_D void st16(void* p, long long i, const float v[16]) {
    if (DT == QT_F32) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            *reinterpret_cast<float4*>((float*)p + i + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            uint32_t u[4];
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (DT == QT_F16) {
                    u[t] = (uint32_t)__half_as_ushort(__float2half_rn(v[8 * q + 2 * t])) |
                           ((uint32_t)__half_as_ushort(__float2half_rn(v[8 * q + 2 * t + 1])) << 16);
                } else {
                    u[t] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[8 * q + 2 * t])) |
                           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[8 * q + 2 * t + 1])) << 16);
                }
            }
            *reinterpret_cast<uint4*>((char*)p + (i + 8 * q) * 2) = make_uint4(u[0], u[1], u[2], u[3]);
        }
    }
}

template <int DT, int LPG, bool DELTA>
__global__ void __launch_bounds__(256) scale_qdq16_kernel(const void* __restrict__ W, const float* __restrict__ s, int N,
                                                          int K, int num_bits, int symmetric, void* __restrict__ out,
                                                          float* __restrict__ delta_f32) {
    const float max_int_s = (float)((1 << (num_bits - 1)) - 1), min_int_s = -(float)(1 << (num_bits - 1));
    const float max_int_a = (float)((1 << num_bits) - 1);
    const int cpr = K >> 4;                                   // 16-column chunks per row (a multiple of LPG)
    const long long total = (long long)N * cpr;
    for (long long base = (long long)blockIdx.x * 256; base < total; base += (long long)gridDim.x * 256) {
        const long long ch = base + threadIdx.x;
        const bool valid = ch < total;                        // whole groups are valid or not (total % LPG == 0)
        const long long chc = valid ? ch : total - 1;
        const long long row = chc / cpr;
        const int c0 = (int)(chc - row * cpr) << 4;
        const long long idx = row * K + c0;
        float w[16], ks[16], keep[16];
        ld16<DT>(W, idx, w);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 a = *reinterpret_cast<const float4*>(s + c0 + 4 * q);
            ks[4 * q] = a.x; ks[4 * q + 1] = a.y; ks[4 * q + 2] = a.z; ks[4 * q + 3] = a.w;
        }
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            keep[i] = rnd<DT>(w[i] * ks[i]);
            mn = fminf(mn, keep[i]);
            mx = fmaxf(mx, keep[i]);
        }
#pragma unroll
        for (int o = LPG / 2; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        float sc, z = 0.f;
        if (symmetric) {
            const float max_val = fmaxf(fmaxf(fabsf(mn), fabsf(mx)), rnd<DT>(1e-5f));
            sc = rnd<DT>(max_val / max_int_s);
        } else {
            sc = rnd<DT>(fmaxf(rnd<DT>(mx - mn), rnd<DT>(1e-5f)) / max_int_a);
            z = -rintf(rnd<DT>(mn / sc));
            z = fminf(fmaxf(z, 0.f), max_int_a);
        }
        float o16[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float dq;
            if (symmetric) {
                float q = rintf(rnd<DT>(keep[i] / sc));
                q = fminf(fmaxf(q, min_int_s), max_int_s);
                dq = rnd<DT>(q * sc);
            } else {
                float q = rnd<DT>(rintf(rnd<DT>(keep[i] / sc)) + z);
                q = fminf(fmaxf(q, 0.f), max_int_a);
                dq = rnd<DT>(rnd<DT>(q - z) * sc);
            }
            o16[i] = dq / ks[i];
        }
        if (!valid) continue;
        if (DELTA) {
            float d[16];
#pragma unroll
            for (int i = 0; i < 16; i++) d[i] = rnd<DT>(o16[i]) - w[i];
            st16<QT_BF16>(out, idx, d);
            st16<QT_F32>(delta_f32, idx, d);
        } else {
            st16<DT>(out, idx, o16);
        }
    }
}

template <int DT, bool DELTA>
static int launch_qdq16(const void* W, const float* s, int N, int K, int gs, int num_bits, int symmetric, void* out,
                        float* delta_f32, cudaStream_t st) {
    const long long total = (long long)N * (K >> 4);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    const unsigned g = (unsigned)blocks;
    if (gs == 128) scale_qdq16_kernel<DT, 8, DELTA><<<g, 256, 0, st>>>(W, s, N, K, num_bits, symmetric, out, delta_f32);
    else if (gs == 64) scale_qdq16_kernel<DT, 4, DELTA><<<g, 256, 0, st>>>(W, s, N, K, num_bits, symmetric, out, delta_f32);
    else scale_qdq16_kernel<DT, 2, DELTA><<<g, 256, 0, st>>>(W, s, N, K, num_bits, symmetric, out, delta_f32);
    return check_launch("awq_scale_qdq16");
}

// sum over i of rnd(a[i] - b[i])^2 ; fp32 per-thread partials, fp64 block total -> atomicAdd(double)
template <int DT>
__global__ void __launch_bounds__(256) sq_err_kernel(const void* __restrict__ a, const void* __restrict__ b,
                                                     long long n, double* __restrict__ out) {
    __shared__ double red[8];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const float d = rnd<DT>(ld<DT>(a, i) - ld<DT>(b, i));
        acc = fmaf(d, d, acc);
    }
    double v = (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int i = 0; i < 8; i++) t += red[i];
        atomicAdd(out, t);
    }
}

}  // namespace awq
}  // namespace qt

using namespace qt;
using namespace qt::awq;

extern "C" {

// colsum [K] fp32 += column sums of the group-normalised |W| [N, K]; gamax_scratch is [N, K/gs] fp32.
// group_size <= 0 means one group per row.  The caller divides by the total row count.
int qt_awq_wmean_accumulate(const void* W, int dtype, int N, int K, int group_size, float* gamax_scratch,
                            float* colsum, void* stream) {
    if (!W || !gamax_scratch || !colsum || N <= 0 || K <= 0) return QT_ERR_INVALID;
    const int gs = group_size > 0 ? group_size : K;
    if (K % gs) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (K + 255) / 256;
    int gy = kNumSMs * 8 / gx + 1;
    if (gy > N) gy = N;
#define QT_RUN(DT)                                                                          \
    group_absmax_kernel<DT><<<(N + 7) / 8, 256, 0, st>>>(W, N, K, gs, gamax_scratch);        \
    wmean_kernel<DT><<<dim3(gx, gy), 256, 0, st>>>(W, gamax_scratch, N, K, gs, colsum);
    switch (dtype) {
        case QT_F32: QT_RUN(QT_F32) break;
        case QT_F16: QT_RUN(QT_F16) break;
        case QT_BF16: QT_RUN(QT_BF16) break;
        default: return QT_ERR_INVALID;
    }
#undef QT_RUN
    int rc = check_launch("awq_wmean");
    return rc ? rc : check_launch("awq_wmean");
}

int qt_awq_scale_qdq(const void* W, int dtype, int N, int K, const float* s, int group_size, int num_bits,
                     int symmetric, void* out, void* stream) {
    if (!W || !s || !out || N <= 0 || K <= 0 || num_bits < 2 || num_bits > 8) return QT_ERR_INVALID;
    const int gs = group_size > 0 ? group_size : K;
    if (K % gs || (gs & 3) || (K & 3) || ((uintptr_t)W & 15) || ((uintptr_t)out & 15) || ((uintptr_t)s & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if ((gs == 32 || gs == 64 || gs == 128) && !(K & 15)) {
        switch (dtype) {
            case QT_F32: return launch_qdq16<QT_F32, false>(W, s, N, K, gs, num_bits, symmetric, out, nullptr, st);
            case QT_F16: return launch_qdq16<QT_F16, false>(W, s, N, K, gs, num_bits, symmetric, out, nullptr, st);
            case QT_BF16: return launch_qdq16<QT_BF16, false>(W, s, N, K, gs, num_bits, symmetric, out, nullptr, st);
            default: return QT_ERR_INVALID;
        }
    }
    const int grid = (N + 7) / 8;
    switch (dtype) {
        case QT_F32: scale_qdq_kernel<QT_F32><<<grid, 256, 0, st>>>(W, s, N, K, gs, num_bits, symmetric, out); break;
        case QT_F16: scale_qdq_kernel<QT_F16><<<grid, 256, 0, st>>>(W, s, N, K, gs, num_bits, symmetric, out); break;
        case QT_BF16: scale_qdq_kernel<QT_BF16><<<grid, 256, 0, st>>>(W, s, N, K, gs, num_bits, symmetric, out); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("awq_scale_qdq");
}

// D = pseudo_quant(W * s) / s - W for the Gram-form loss of a single-Linear parent (qt_awq_gram_loss): the same
// candidate weight as qt_awq_scale_qdq (rounded to the weight's dtype, as torch would hold it), minus W, written as
// bf16 [N, K] (delta_bf16) and fp32 [N, K] (delta_f32).  group_size 32 / 64 / 128 only.
int qt_awq_scale_qdq_delta(const void* W, int dtype, int N, int K, const float* s, int group_size, int num_bits,
                           int symmetric, void* delta_bf16, float* delta_f32, void* stream) {
    if (!W || !s || !delta_bf16 || !delta_f32 || N <= 0 || K <= 0 || num_bits < 2 || num_bits > 8) return QT_ERR_INVALID;
    if (!(group_size == 32 || group_size == 64 || group_size == 128) || (K % group_size)) return QT_ERR_UNSUPPORTED;
    if (((uintptr_t)W | (uintptr_t)delta_bf16 | (uintptr_t)delta_f32 | (uintptr_t)s) & 15) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case QT_F32: return launch_qdq16<QT_F32, true>(W, s, N, K, group_size, num_bits, symmetric, delta_bf16, delta_f32, st);
        case QT_F16: return launch_qdq16<QT_F16, true>(W, s, N, K, group_size, num_bits, symmetric, delta_bf16, delta_f32, st);
        case QT_BF16: return launch_qdq16<QT_BF16, true>(W, s, N, K, group_size, num_bits, symmetric, delta_bf16, delta_f32, st);
        default: return QT_ERR_INVALID;
    }
}

// *out (device double, zeroed by the caller) += sum (a - b)^2 over n elements
int qt_sq_err_sum(const void* a, const void* b, int dtype, int64_t n, double* out, void* stream) {
    if (!a || !b || !out || n < 0) return QT_ERR_INVALID;
    if (n == 0) return QT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n + 255) / 256;
    const long long cap = (long long)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    switch (dtype) {
        case QT_F32: sq_err_kernel<QT_F32><<<(unsigned)blocks, 256, 0, st>>>(a, b, n, out); break;
        case QT_F16: sq_err_kernel<QT_F16><<<(unsigned)blocks, 256, 0, st>>>(a, b, n, out); break;
        case QT_BF16: sq_err_kernel<QT_BF16><<<(unsigned)blocks, 256, 0, st>>>(a, b, n, out); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("sq_err_sum");
}

}  // extern "C"
