// Shared helpers for the sm_100a kernels behind the C-ABI (include/quantool_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define QT_OK 0
#define QT_ERR_INVALID (-1)      // bad argument (shape, alignment, enum)
#define QT_ERR_CUDA (-2)         // a CUDA runtime call or launch failed; see qt_last_error()
#define QT_ERR_UNSUPPORTED (-3)  // valid request that this build does not implement

// element types accepted at the boundary
#define QT_F32 0
#define QT_F16 1
#define QT_BF16 2

#define QT_HD __host__ __device__ __forceinline__
#define QT_D __device__ __forceinline__

namespace qt {

void set_last_error(const char* what, cudaError_t e);
int check_launch(const char* what);
unsigned long long launch_count();

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

QT_D float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
QT_D float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
QT_D float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit streaming load / store (read-once data: do not pollute L1)
QT_D uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
QT_D void stg_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

QT_D float bf16_bits_to_float(uint32_t b16) { return __uint_as_float(b16 << 16); }
QT_D float f16_bits_to_float(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }

// x ~= hi + lo with hi = tf32(x) and lo = tf32(x - hi), both rounded to nearest (low 13 mantissa bits cleared): the
// operand form of the 3xTF32 tensor-core GEMMs (tgemm.cu; splits in split.cu).  x - hi is exact in fp32; rounding it here
// (instead of letting the tensor core truncate the low bits when it reads the operand) halves the operand error and
// removes its bias: |x - hi - lo| <= 2^-22 |x|.
QT_D float tf32_rn(float e) {
    const uint32_t b = __float_as_uint(e);
    return ((b & 0x7F800000u) == 0x7F800000u) ? e : __uint_as_float((b + 0x1000u) & 0xFFFFE000u);
}
QT_D void tf32_split(float e, float& h, float& l) {
    h = tf32_rn(e);
    l = tf32_rn(e - h);
}

// fp32 -> fp16 (RNE) -> fp32: the reference's f16-GGUF intermediate
QT_HD float round_via_f16(float v) { return __half2float(__float2half_rn(v)); }

// Load 8 consecutive elements starting at element index 8*idx8 as fp32.
template <int DT>
QT_D void load8(const void* __restrict__ src, int64_t idx8, float v[8]) {
    if (DT == QT_F32) {
        const uint4 a = ldg_stream((const char*)src + idx8 * 32);
        const uint4 b = ldg_stream((const char*)src + idx8 * 32 + 16);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y);
        v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y);
        v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    } else {
        const uint4 a = ldg_stream((const char*)src + idx8 * 16);
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

}  // namespace qt
