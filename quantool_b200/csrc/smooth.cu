// SmoothQuant / AWQ bandwidth-bound reductions and folds on sm_100a.
//
// Replaces UPSTREAM llmcompressor SmoothQuantModifier (`_calculate_smoothing_scales`,
// `_apply_smoothing`, the min/max forward hook) reached from
// ref/src/quantool/methods/llm_compressor/smoothquant/smoothquant.py:77-84, and the per-channel
// statistics of AWQModifier (`_accumulate_mean`) reached from
// ref/src/quantool/methods/llm_compressor/awq/awq.py:81.
//
// Everything here is an HBM streaming pass: a column reduction over [T, K] activations
// (128-bit loads, 8 columns per thread, partial results combined with order-insensitive
// atomics), or an elementwise fold W *= s / W /= s evaluated in the tensor's own dtype.
#include <math.h>

#include "common.cuh"

namespace qt {
namespace smooth {

template <int DT>
QT_D float rnd(float v) {
    if (DT == QT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (DT == QT_F16) return __half2float(__float2half_rn(v));
    return v;
}

// float atomic min/max through the sign-split integer ordering trick
QT_D void atomic_max_f(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
QT_D void atomic_min_f(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

enum { OP_MINMAX = 0, OP_ABSSUM = 1 };

// X: [T, K] (K % 8 == 0).  OP_MINMAX: out0[c] = min(out0[c], min_t X), out1[c] = max(...)
//                           OP_ABSSUM: out0[c] += sum_t |X|   (fp32)
// block = 32 column-groups (8 columns each) x 8 row lanes; grid = (ceil(K/256), row blocks)
template <int DT, int OP>
__global__ void __launch_bounds__(256) col_reduce_kernel(const void* __restrict__ X, long long T, int K,
                                                         float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ float sa[8][256 + 8];
    __shared__ float sb[8][256 + 8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + tx * 8;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = (OP == OP_MINMAX) ? 3.402823466e+38f : 0.f; b[i] = -3.402823466e+38f; }
    if (c0 < K) {
        for (long long r = (long long)blockIdx.y * 8 + ty; r < T; r += (long long)gridDim.y * 8) {
            float v[8];
            if (DT == QT_F32) {
                const float4 p = *reinterpret_cast<const float4*>((const float*)X + r * K + c0);
                const float4 q = *reinterpret_cast<const float4*>((const float*)X + r * K + c0 + 4);
                v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
            } else {
                load8<DT>(X, (r * K + c0) / 8, v);
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (OP == OP_MINMAX) { a[i] = fminf(a[i], v[i]); b[i] = fmaxf(b[i], v[i]); }
                else a[i] += fabsf(v[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) { sa[ty][tx * 8 + i] = a[i]; if (OP == OP_MINMAX) sb[ty][tx * 8 + i] = b[i]; }
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < K) {
        float ra = sa[0][threadIdx.x], rb = (OP == OP_MINMAX) ? sb[0][threadIdx.x] : 0.f;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            if (OP == OP_MINMAX) { ra = fminf(ra, sa[j][threadIdx.x]); rb = fmaxf(rb, sb[j][threadIdx.x]); }
            else ra += sa[j][threadIdx.x];
        }
        if (OP == OP_MINMAX) { atomic_min_f(out0 + c, ra); atomic_max_f(out1 + c, rb); }
        else atomicAdd(out0 + c, ra);
    }
}

__global__ void fill_kernel(float* p, int n, float v) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) p[i] = v;
}

// SmoothQuant scales in compute dtype CD:
//   act = max - min ; w = 2 * wabsmax ; s = act^a / w^(1-a) ; s = where(w > 0, s, act) ; s = max(s, 1e-5)
template <int CD>
__global__ void smooth_scales_kernel(const float* __restrict__ amin, const float* __restrict__ amax,
                                     const float* __restrict__ wmin, const float* __restrict__ wmax, float alpha,
                                     float one_minus_alpha, float* __restrict__ s_out, int K) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= K) return;
    const float act = rnd<CD>(amax[c] - amin[c]);
    const float wabs = fmaxf(fabsf(wmin[c]), fabsf(wmax[c]));
    const float w = rnd<CD>(2.0f * wabs);
    // torch casts a python-scalar exponent to the tensor dtype: bf16 tensors see pow(x, bf16(alpha))
    const float num = rnd<CD>((float)pow((double)act, (double)rnd<CD>(alpha)));
    const float den = rnd<CD>((float)pow((double)w, (double)rnd<CD>(one_minus_alpha)));
    float s = rnd<CD>(num / den);
    s = (w > 0.f) ? s : act;
    s = fmaxf(s, 1e-5f);   // torch.maximum(scales, torch.Tensor([1e-5])): fp32 floor, result promoted to fp32
    s_out[c] = s;
}

// W[n][c] (op)= s[c]  (by_row: s[n]) evaluated in W's dtype; s is fp32.  One thread = 8 consecutive
// columns (one 128-bit load + store for 16-bit types); K % 8 == 0.
template <int DT, bool DIV, bool BY_ROW>
__global__ void __launch_bounds__(256) scale_kernel(void* __restrict__ W, const float* __restrict__ s, int N, int K) {
    const int c8 = (blockIdx.x * 256 + threadIdx.x) * 8;
    const int n = blockIdx.y;
    if (c8 >= K) return;
    const long long idx = (long long)n * K + c8;
    float v[8], sv[8];
#pragma unroll
    for (int i = 0; i < 8; i++) sv[i] = BY_ROW ? s[n] : s[c8 + i];
    if (DT == QT_F32) {
        const float4 a = *reinterpret_cast<const float4*>((float*)W + idx), b = *reinterpret_cast<const float4*>((float*)W + idx + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 a = *reinterpret_cast<const uint4*>((char*)W + idx * 2);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (DT == QT_F16) { v[2 * i] = f16_bits_to_float(w[i] & 0xffffu); v[2 * i + 1] = f16_bits_to_float(w[i] >> 16); }
            else { v[2 * i] = bf16_bits_to_float(w[i] & 0xffffu); v[2 * i + 1] = bf16_bits_to_float(w[i] >> 16); }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = DIV ? v[i] / sv[i] : v[i] * sv[i];
    if (DT == QT_F32) {
        *reinterpret_cast<float4*>((float*)W + idx) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>((float*)W + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (DT == QT_F16)
                w[i] = (uint32_t)__half_as_ushort(__float2half_rn(v[2 * i])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[2 * i + 1])) << 16);
            else
                w[i] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * i])) |
                       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * i + 1])) << 16);
        }
        *reinterpret_cast<uint4*>((char*)W + idx * 2) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int DT, int OP>
static void launch_col_reduce(const void* X, long long T, int K, float* o0, float* o1, cudaStream_t st) {
    const int gx = (K + 255) / 256;
    long long gy = (T + 7) / 8;
    const long long cap = (long long)kNumSMs * 8 / gx + 1;   // ~8 CTAs per SM in flight
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    col_reduce_kernel<DT, OP><<<dim3(gx, (unsigned)gy), 256, 0, st>>>(X, T, K, o0, o1);
}

}  // namespace smooth
}  // namespace qt

using namespace qt;
using namespace qt::smooth;

extern "C" {

int qt_fill_f32(float* p, int n, float v, void* stream) {
    if (!p || n < 0) return QT_ERR_INVALID;
    if (n == 0) return QT_OK;
    fill_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, n, v);
    return check_launch("fill_f32");
}

// running per-channel min / max over the rows of X [T, K]; mn/mx are fp32 [K], initialised by the
// caller to +FLT_MAX / -FLT_MAX (qt_fill_f32) before the first batch.
int qt_channel_minmax(const void* X, int dtype, int64_t T, int K, float* mn, float* mx, void* stream) {
    if (!X || !mn || !mx || T < 0 || K <= 0 || (K & 7) || ((uintptr_t)X & 15)) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case QT_F32: launch_col_reduce<QT_F32, OP_MINMAX>(X, T, K, mn, mx, st); break;
        case QT_F16: launch_col_reduce<QT_F16, OP_MINMAX>(X, T, K, mn, mx, st); break;
        case QT_BF16: launch_col_reduce<QT_BF16, OP_MINMAX>(X, T, K, mn, mx, st); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("channel_minmax");
}

// sum[c] += sum_t |X[t][c]| (fp32); the caller divides by the token count (AWQ x_mean)
int qt_channel_abs_sum(const void* X, int dtype, int64_t T, int K, float* sum, void* stream) {
    if (!X || !sum || T < 0 || K <= 0 || (K & 7) || ((uintptr_t)X & 15)) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case QT_F32: launch_col_reduce<QT_F32, OP_ABSSUM>(X, T, K, sum, nullptr, st); break;
        case QT_F16: launch_col_reduce<QT_F16, OP_ABSSUM>(X, T, K, sum, nullptr, st); break;
        case QT_BF16: launch_col_reduce<QT_BF16, OP_ABSSUM>(X, T, K, sum, nullptr, st); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("channel_abs_sum");
}

int qt_smooth_scales(const float* amin, const float* amax, const float* wmin, const float* wmax, float alpha,
                     float one_minus_alpha, int compute_dtype, float* s_out, int K, void* stream) {
    if (!amin || !amax || !wmin || !wmax || !s_out || K <= 0) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = (K + 255) / 256;
    switch (compute_dtype) {
        case QT_F32: smooth_scales_kernel<QT_F32><<<g, 256, 0, st>>>(amin, amax, wmin, wmax, alpha, one_minus_alpha, s_out, K); break;
        case QT_F16: smooth_scales_kernel<QT_F16><<<g, 256, 0, st>>>(amin, amax, wmin, wmax, alpha, one_minus_alpha, s_out, K); break;
        case QT_BF16: smooth_scales_kernel<QT_BF16><<<g, 256, 0, st>>>(amin, amax, wmin, wmax, alpha, one_minus_alpha, s_out, K); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("smooth_scales");
}

// in place: W[n][c] *= s[c] (divide != 0: /=), or by_row != 0: W[n][c] (op)= s[n]
int qt_scale_matrix(void* W, int dtype, int N, int K, const float* s, int divide, int by_row, void* stream) {
    if (!W || !s || N <= 0 || K <= 0 || N > 65535 || (K & 7) || ((uintptr_t)W & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((K / 8 + 255) / 256, N);
#define QT_SCALE(DT)                                                                                   \
    if (divide) { if (by_row) scale_kernel<DT, true, true><<<grid, 256, 0, st>>>(W, s, N, K);          \
                  else scale_kernel<DT, true, false><<<grid, 256, 0, st>>>(W, s, N, K); }              \
    else        { if (by_row) scale_kernel<DT, false, true><<<grid, 256, 0, st>>>(W, s, N, K);         \
                  else scale_kernel<DT, false, false><<<grid, 256, 0, st>>>(W, s, N, K); }
    switch (dtype) {
        case QT_F32: QT_SCALE(QT_F32) break;
        case QT_F16: QT_SCALE(QT_F16) break;
        case QT_BF16: QT_SCALE(QT_BF16) break;
        default: return QT_ERR_INVALID;
    }
#undef QT_SCALE
    return check_launch("scale_matrix");
}

}  // extern "C"
