// Round-to-nearest quantization, qparams observer and int32 packing on sm_100a: the
// compressed-tensors primitives the reference's llm-compressor plugins end in
// (ref/src/quantool/methods/llm_compressor/base.py:188 `save_pretrained(save_compressed=True)`).
//   minmax observer + calculate_qparams   CT/quantization/utils/helpers.py:50-137   (row a4)
//   quantize -> int8 codes                CT/quantization/lifecycle/forward.py:36-73 (row a6/a7)
//   pack_to_int32                         CT/compressors/pack_quantized/helpers.py:20-89 (row a6)
//
// torch evaluates these in the tensor's dtype: for bf16/fp16 tensors every elementwise op is
// "compute in fp32, round the result to the storage dtype".  CD (compute dtype) reproduces
// that rounding after every op so the integer codes match the reference's artifact exactly.
// All kernels are HBM streaming kernels: one warp walks one row with 128-bit loads.
#include "quant_math.cuh"

namespace qt {
namespace quant {

template <int CD>
QT_D float rnd(float v) {
    if (CD == QT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (CD == QT_F16) return __half2float(__float2half_rn(v));
    return v;
}
template <int CD>
QT_D float eps_of() {
    return CD == QT_BF16 ? 0.0078125f : CD == QT_F16 ? 0.0009765625f : 1.1920928955078125e-07f;
}

template <int DT>
QT_D float ld_elem(const void* p, long long i) {
    if (DT == QT_F32) return reinterpret_cast<const float*>(p)[i];
    if (DT == QT_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <int DT>
QT_D void st_elem(void* p, long long i, float v) {
    if (DT == QT_F32) reinterpret_cast<float*>(p)[i] = v;
    else if (DT == QT_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// calculate_qparams evaluated in compute dtype CD
template <int CD>
QT_D void calc_qparams_cd(float mn, float mx, int num_bits, bool symmetric, float& scale, float& zp) {
    mn = fminf(mn, 0.f);
    mx = fmaxf(mx, 0.f);
    const QRange r = int_range(num_bits);
    const float bit_range = r.qmax - r.qmin;
    if (symmetric) {
        const float max_val_pos = fmaxf(fabsf(mn), fabsf(mx));
        scale = rnd<CD>(max_val_pos / (bit_range / 2.f));
        zp = 0.f;
    } else {
        scale = rnd<CD>(rnd<CD>(mx - mn) / bit_range);
        float z = rnd<CD>(r.qmin - rnd<CD>(mn / scale));
        z = fminf(fmaxf(z, r.qmin), r.qmax);
        z = fminf(fmaxf(z, -128.f), 127.f);
        z = rintf(z);
        zp = (z == z) ? z : 0.f;
    }
    if (scale == 0.f) scale = eps_of<CD>();
}

// clamp(round(x/scale + zp)) in compute dtype CD -> integer code (as float)
template <int CD>
QT_D float quant_code(float x, float scale, float zp, QRange r) {
    float v = rnd<CD>(x / scale);
    v = rnd<CD>(v + zp);
    v = fminf(fmaxf(v, r.qmin), r.qmax);
    return rintf(v);
}

// ---- minmax observer: one warp per row, loop over groups -------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) minmax_qparams_kernel(const void* __restrict__ W, int N, int K, int group_size,
                                                             int num_bits, int symmetric, float* __restrict__ scale,
                                                             float* __restrict__ zp) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= N) return;
    const int gs = group_size > 0 ? group_size : K;
    const int G = K / gs;
    for (int g = 0; g < G; g++) {
        float mn = 3.402823466e+38f, mx = -3.402823466e+38f;
        for (int c = lane; c < gs; c += 32) {
            const float v = ld_elem<DT>(W, (long long)row * K + (long long)g * gs + c);
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        mn = warp_min(mn);
        mx = warp_max(mx);
        float s, z;
        calc_qparams_cd<DT>(mn, mx, num_bits, symmetric != 0, s, z);
        if (lane == 0) {
            scale[(long long)row * G + g] = s;
            zp[(long long)row * G + g] = z;
        }
    }
}

// ---- codes (+ optional fake-quant output) ----------------------------------------------------
// code[n][c] = quant(W[n][c], scale[n][grp(c)], zp[n][grp(c)]), grp(c) = g_idx ? g_idx[c] : c/gs
// scale is given in the same dtype as W (the artifact's weight_scale), zp as fp32 integers or null.
template <int DT>
__global__ void __launch_bounds__(256) quantize_codes_kernel(const void* __restrict__ W, const void* __restrict__ scale,
                                                             const float* __restrict__ zp, const int* __restrict__ g_idx,
                                                             int N, int K, int G, int group_size, int num_bits,
                                                             int8_t* __restrict__ codes, void* __restrict__ dq_out) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (c >= K) return;
    const int gs = group_size > 0 ? group_size : K;
    const int g = g_idx ? g_idx[c] : c / gs;
    const float s = ld_elem<DT>(scale, (long long)n * G + g);
    const float z = zp ? zp[(long long)n * G + g] : 0.f;
    const float x = ld_elem<DT>(W, (long long)n * K + c);
    const QRange r = int_range(num_bits);
    const float q = quant_code<DT>(x, s, z, r);
    if (codes) codes[(long long)n * K + c] = (int8_t)(int)q;
    if (dq_out) {
        // _dequantize: (q - zp) * scale in scale.dtype
        const float dq = rnd<DT>(rnd<DT>(q - z) * s);
        st_elem<DT>(dq_out, (long long)n * K + c, dq);
    }
}

// ---- pack_to_int32 along dim 1: (q + 2^(b-1)) as unsigned field i at bit b*i, zero padded ------
__global__ void __launch_bounds__(256) pack_int32_kernel(const int8_t* __restrict__ codes, int N, int K, int num_bits,
                                                         int32_t* __restrict__ packed, int Kp) {
    const int p = blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (p >= Kp) return;
    const int pf = 32 / num_bits;
    const int offset = 1 << (num_bits - 1);
    uint32_t word = 0;
    for (int i = 0; i < pf; i++) {
        const int c = p * pf + i;
        if (c < K) {
            const uint32_t u = (uint32_t)((int)codes[(long long)n * K + c] + offset) & 0xffu;
            word |= u << (num_bits * i);
        }
    }
    packed[(long long)n * Kp + p] = (int32_t)word;
}

// unpack_from_int32 (for round-trip checks and the AutoGPTQ/AutoAWQ views)
__global__ void __launch_bounds__(256) unpack_int32_kernel(const int32_t* __restrict__ packed, int N, int K, int num_bits,
                                                           int8_t* __restrict__ codes, int Kp) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int n = blockIdx.y;
    if (c >= K) return;
    const int pf = 32 / num_bits;
    const uint32_t word = (uint32_t)packed[(long long)n * Kp + c / pf];
    const uint32_t mask = (1u << num_bits) - 1u;
    const int u = (int)((word >> (num_bits * (c % pf))) & mask);
    codes[(long long)n * K + c] = (int8_t)(u - (1 << (num_bits - 1)));
}

template <template <int> class F, class... Args>
static int by_dtype(int dt, Args... args) {
    switch (dt) {
        case QT_F32: F<QT_F32>::run(args...); return QT_OK;
        case QT_F16: F<QT_F16>::run(args...); return QT_OK;
        case QT_BF16: F<QT_BF16>::run(args...); return QT_OK;
    }
    return QT_ERR_INVALID;
}

template <int DT>
struct RunMinmax {
    static void run(const void* W, int N, int K, int gs, int nb, int sym, float* s, float* z, cudaStream_t st) {
        minmax_qparams_kernel<DT><<<(N + 7) / 8, 256, 0, st>>>(W, N, K, gs, nb, sym, s, z);
    }
};
template <int DT>
struct RunCodes {
    static void run(const void* W, const void* s, const float* z, const int* gi, int N, int K, int G, int gs, int nb,
                    int8_t* codes, void* dq, cudaStream_t st) {
        dim3 grid((K + 255) / 256, N);
        quantize_codes_kernel<DT><<<grid, 256, 0, st>>>(W, s, z, gi, N, K, G, gs, nb, codes, dq);
    }
};

}  // namespace quant
}  // namespace qt

using namespace qt;
using namespace qt::quant;

extern "C" {

int qt_minmax_qparams(const void* W, int dtype, int N, int K, int group_size, int num_bits, int symmetric,
                      float* scale, float* zp, void* stream) {
    if (!W || !scale || !zp || N <= 0 || K <= 0) return QT_ERR_INVALID;
    if (group_size > 0 && K % group_size) return QT_ERR_INVALID;
    int rc = by_dtype<RunMinmax>(dtype, W, N, K, group_size, num_bits, symmetric, scale, zp, (cudaStream_t)stream);
    return rc ? rc : check_launch("minmax_qparams");
}

int qt_quantize_codes(const void* W, const void* scale, const float* zp, const int* g_idx, int dtype, int N, int K,
                      int G, int group_size, int num_bits, int8_t* codes, void* dq_out, void* stream) {
    if (!W || !scale || N <= 0 || K <= 0 || G <= 0 || (!codes && !dq_out)) return QT_ERR_INVALID;
    int rc = by_dtype<RunCodes>(dtype, W, scale, zp, g_idx, N, K, G, group_size, num_bits, codes, dq_out,
                                (cudaStream_t)stream);
    return rc ? rc : check_launch("quantize_codes");
}

int qt_pack_int32(const int8_t* codes, int N, int K, int num_bits, int32_t* packed, void* stream) {
    if (!codes || !packed || N <= 0 || K <= 0 || num_bits < 1 || num_bits > 8) return QT_ERR_INVALID;
    const int pf = 32 / num_bits, Kp = (K + pf - 1) / pf;
    dim3 grid((Kp + 255) / 256, N);
    pack_int32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(codes, N, K, num_bits, packed, Kp);
    return check_launch("pack_int32");
}

int qt_unpack_int32(const int32_t* packed, int N, int K, int num_bits, int8_t* codes, void* stream) {
    if (!codes || !packed || N <= 0 || K <= 0 || num_bits < 1 || num_bits > 8) return QT_ERR_INVALID;
    const int pf = 32 / num_bits, Kp = (K + pf - 1) / pf;
    dim3 grid((K + 255) / 256, N);
    unpack_int32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, N, K, num_bits, codes, Kp);
    return check_launch("unpack_int32");
}

}  // extern "C"
