// Elementwise pieces of the calibration forward of a Llama decoder layer on sm_100a.
//
// The reference runs the calibration data through the HF model inside `oneshot`
// (ref/src/quantool/methods/llm_compressor/base.py:162, llm-compressor's sequential pipeline):
// every layer is run twice per round (statistics pass with the original weights, propagation pass
// with the quantized ones).  The GEMMs and attention of that forward are library calls; the
// normalisation, rotary embedding and gated activation between them were ~25 % of the forward as
// separate torch kernels (each op a full read+write of a [T, 4096..14336] tensor).  These three
// kernels do each of them in ONE pass over HBM with the same rounding points as the HF modules
// (transformers LlamaRMSNorm / apply_rotary_pos_emb / LlamaMLP): every torch op of the original
// expression is one rounding to the tensor dtype here.
//
// HBM bound.  Algorithmic bytes per element (bf16): rms_norm 2 read + 2 write (the second read of
// the row hits L2), rope 2 + 2 in place, silu_mul 4 read + 2 write.
#include <math.h>

#include "common.cuh"

namespace qt {
namespace fwd {

template <int DT>
QT_D float rnd(float v) {
    if (DT == QT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    if (DT == QT_F16) return __half2float(__float2half_rn(v));
    return v;
}

template <int DT>
QT_D uint32_t pack2(float a, float b) {
    if (DT == QT_BF16) {
        const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<const uint32_t*>(&p);
    }
    const __half2 p = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&p);
}

template <int DT>
QT_D void store8(void* dst, int64_t idx8, const float v[8]) {
    if (DT == QT_F32) {
        float4* p = reinterpret_cast<float4*>((char*)dst + idx8 * 32);
        p[0] = make_float4(v[0], v[1], v[2], v[3]);
        p[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
        uint4 o;
        o.x = pack2<DT>(v[0], v[1]); o.y = pack2<DT>(v[2], v[3]);
        o.z = pack2<DT>(v[4], v[5]); o.w = pack2<DT>(v[6], v[7]);
        *reinterpret_cast<uint4*>((char*)dst + idx8 * 16) = o;
    }
}

// plain (cached) 8-element load: rows are re-read, so no streaming hint here
template <int DT>
QT_D void load8c(const void* src, int64_t idx8, float v[8]) {
    if (DT == QT_F32) {
        const float4* p = reinterpret_cast<const float4*>((const char*)src + idx8 * 32);
        const float4 a = p[0], b = p[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        const uint4 a = *reinterpret_cast<const uint4*>((const char*)src + idx8 * 16);
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (DT == QT_F16) {
                v[2 * i] = f16_bits_to_float(w[i] & 0xffffu);
                v[2 * i + 1] = f16_bits_to_float(w[i] >> 16);
            } else {
                v[2 * i] = bf16_bits_to_float(w[i] & 0xffffu);
                v[2 * i + 1] = bf16_bits_to_float(w[i] >> 16);
            }
        }
    }
}

// out[t] = w * (x[t].float() * rsqrt(mean(x[t]^2) + eps)).to(dtype)      one CTA of 128 threads per row
template <int DT>
__global__ void __launch_bounds__(128) rms_norm_kernel(const void* __restrict__ X, const void* __restrict__ Wt,
                                                       void* __restrict__ Out, int H, float eps) {
    __shared__ float red[4];
    const int64_t row8 = (int64_t)blockIdx.x * (H / 8);
    float ss = 0.f;
    for (int c = threadIdx.x; c < H / 8; c += 128) {
        float v[8];
        load8c<DT>(X, row8 + c, v);
#pragma unroll
        for (int i = 0; i < 8; i++) ss += v[i] * v[i];
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    const float tot = (red[0] + red[1]) + (red[2] + red[3]);
    const float r = rsqrtf(tot / (float)H + eps);
    for (int c = threadIdx.x; c < H / 8; c += 128) {
        float v[8], w[8], o[8];
        load8c<DT>(X, row8 + c, v);
        load8c<DT>(Wt, c, w);
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = rnd<DT>(w[i] * rnd<DT>(v[i] * r));
        store8<DT>(Out, row8 + c, o);
    }
}

// In place on X[T, n_heads*hd] (the projection output, token-major): x = x*cos + rotate_half(x)*sin
// with cos/sin [S, hd] in the same dtype and position = t % S.  One thread per 8 elements of the
// first half of a head and their partners in the second half.
template <int DT>
__global__ void __launch_bounds__(256) rope_kernel(void* __restrict__ X, const void* __restrict__ Cos,
                                                   const void* __restrict__ Sin, int64_t T, int S, int n_heads,
                                                   int hd) {
    const int half8 = hd / 16;                       // 8-element chunks per half head
    const int64_t gid = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int64_t total = T * n_heads * half8;
    if (gid >= total) return;
    const int c = (int)(gid % half8);
    const int64_t th = gid / half8;                  // token * n_heads + head
    const int64_t t = th / n_heads;
    const int pos = (int)(t % S);
    const int64_t x1 = th * (hd / 8) + c, x2 = x1 + half8;
    const int64_t t1 = (int64_t)pos * (hd / 8) + c, t2 = t1 + half8;
    float a[8], b[8], c1[8], s1[8], c2[8], s2[8], o1[8], o2[8];
    load8c<DT>(X, x1, a); load8c<DT>(X, x2, b);
    load8c<DT>(Cos, t1, c1); load8c<DT>(Sin, t1, s1);
    load8c<DT>(Cos, t2, c2); load8c<DT>(Sin, t2, s2);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        o1[i] = rnd<DT>(rnd<DT>(a[i] * c1[i]) + rnd<DT>(-b[i] * s1[i]));
        o2[i] = rnd<DT>(rnd<DT>(b[i] * c2[i]) + rnd<DT>(a[i] * s2[i]));
    }
    store8<DT>(X, x1, o1);
    store8<DT>(X, x2, o2);
}

// out = silu(gate) * up, both products rounded to the tensor dtype like the two torch ops
template <int DT>
__global__ void __launch_bounds__(256) silu_mul_kernel(const void* __restrict__ G, const void* __restrict__ U,
                                                       void* __restrict__ Out, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        float g[8], u[8], o[8];
        load8<DT>(G, i, g);
        load8<DT>(U, i, u);
#pragma unroll
        for (int k = 0; k < 8; k++) o[k] = rnd<DT>(rnd<DT>(g[k] / (1.0f + expf(-g[k]))) * u[k]);
        store8<DT>(Out, i, o);
    }
}

template <template <int> class Launch, class... A>
static int by_dtype(int dt, A... a) {
    switch (dt) {
        case QT_F32: Launch<QT_F32>::run(a...); return QT_OK;
        case QT_F16: Launch<QT_F16>::run(a...); return QT_OK;
        case QT_BF16: Launch<QT_BF16>::run(a...); return QT_OK;
    }
    return QT_ERR_INVALID;
}

template <int DT> struct LaunchNorm {
    static void run(const void* X, const void* W, void* O, int64_t T, int H, float eps, cudaStream_t st) {
        rms_norm_kernel<DT><<<(unsigned)T, 128, 0, st>>>(X, W, O, H, eps);
    }
};
template <int DT> struct LaunchRope {
    static void run(void* X, const void* C, const void* S, int64_t T, int seq, int nh, int hd, cudaStream_t st) {
        const int64_t total = T * nh * (hd / 16);
        rope_kernel<DT><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(X, C, S, T, seq, nh, hd);
    }
};
template <int DT> struct LaunchSilu {
    static void run(const void* G, const void* U, void* O, int64_t n8, cudaStream_t st) {
        const int64_t want = (n8 + 255) / 256;
        const int64_t cap = (int64_t)kNumSMs * 16;
        silu_mul_kernel<DT><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(G, U, O, n8);
    }
};

}  // namespace fwd
}  // namespace qt

using namespace qt;

extern "C" {

int qt_rms_norm(const void* x, const void* weight, void* out, int dtype, int64_t T, int H, float eps, void* stream) {
    if (T < 0 || H <= 0 || (H & 7)) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;
    if (!x || !weight || !out || T > 2147483647LL) return QT_ERR_INVALID;
    const int rc = fwd::by_dtype<fwd::LaunchNorm>(dtype, x, weight, out, T, H, eps, (cudaStream_t)stream);
    return rc ? rc : check_launch("qt_rms_norm");
}

int qt_rope_inplace(void* x, const void* cos_t, const void* sin_t, int dtype, int64_t T, int seq, int n_heads,
                    int head_dim, void* stream) {
    if (T < 0 || seq <= 0 || n_heads <= 0 || head_dim <= 0 || (head_dim & 15)) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;
    if (!x || !cos_t || !sin_t) return QT_ERR_INVALID;
    const int rc = fwd::by_dtype<fwd::LaunchRope>(dtype, x, cos_t, sin_t, T, seq, n_heads, head_dim, (cudaStream_t)stream);
    return rc ? rc : check_launch("qt_rope_inplace");
}

int qt_silu_mul(const void* gate, const void* up, void* out, int dtype, int64_t n, void* stream) {
    if (n < 0 || (n & 7)) return QT_ERR_INVALID;
    if (n == 0) return QT_OK;
    if (!gate || !up || !out) return QT_ERR_INVALID;
    const int rc = fwd::by_dtype<fwd::LaunchSilu>(dtype, gate, up, out, n / 8, (cudaStream_t)stream);
    return rc ? rc : check_launch("qt_silu_mul");
}

}  // extern "C"
