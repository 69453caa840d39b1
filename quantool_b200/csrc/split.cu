// tf32 operand splits for the 3xTF32 tensor-core GEMMs (tgemm.cu): x ~= hi + lo with hi = tf32(x),
// lo = tf32(x - hi), both round-to-nearest (common.cuh tf32_split).  Used by the inverse-Hessian chain and by the
// lazy-batch update W[:, i2:] -= Err * U[i1:i2, i2:] of UPSTREAM gptq_quantize.py quantize_weight (SURVEY.md
// section A.4, row a5), whose B operand is the transposed split of U.
#include "common.cuh"

namespace qt {
namespace split {

QT_D void split1(float e, float& h, float& l) { tf32_split(e, h, l); }

// hi = tf32(x), lo = tf32(x - hi), both round to nearest (low 13 mantissa bits cleared)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi,
                                                         float* __restrict__ lo, long long n4) {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
        const float4 v = reinterpret_cast<const float4*>(x)[i];
        float4 h, l;
        split1(v.x, h.x, l.x); split1(v.y, h.y, l.y); split1(v.z, h.z, l.z); split1(v.w, h.w, l.w);
        reinterpret_cast<float4*>(hi)[i] = h;
        reinterpret_cast<float4*>(lo)[i] = l;
    }
}

// hi_t[j][i] / lo_t[j][i] = split(U[i][j]) for j >= i (upper-triangular U), 0 elsewhere: the
// transposed, tf32-split copy of U that the lazy GEMM reads as a K-major B operand
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ U, float* __restrict__ hi_t,
                                                              float* __restrict__ lo_t, int K) {
    __shared__ float t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        t[r][tx] = (i < K && j < K) ? U[(long long)i * K + j] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bj * 32 + r, i = bi * 32 + tx;   // output row j, column i
        if (j < K && i < K) {
            float h, l;
            split1(t[tx][r], h, l);
            hi_t[(long long)j * K + i] = h;
            lo_t[(long long)j * K + i] = l;
        }
    }
}

}  // namespace split
}  // namespace qt

using namespace qt;

extern "C" {

// hi = tf32(x), lo = tf32(x - hi); n % 4 == 0, 16-byte aligned pointers
int qt_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
    if (!x || !hi || !lo || n < 0 || (n & 3)) return QT_ERR_INVALID;
    if (((uintptr_t)x & 15) || ((uintptr_t)hi & 15) || ((uintptr_t)lo & 15)) return QT_ERR_INVALID;
    if (n == 0) return QT_OK;
    long long n4 = n / 4, blocks = (n4 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    split::split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n4);
    return check_launch("split_tf32");
}

// ut_hi/ut_lo [K,K] = transposed tf32 split of U [K,K] (B operand of the lazy update)
int qt_split_tf32_transpose(const float* U, float* ut_hi, float* ut_lo, int K, void* stream) {
    if (!U || !ut_hi || !ut_lo || K <= 0) return QT_ERR_INVALID;
    dim3 grid((K + 31) / 32, (K + 31) / 32);
    split::split_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(U, ut_hi, ut_lo, K);
    return check_launch("split_tf32_transpose");
}

}  // extern "C"
