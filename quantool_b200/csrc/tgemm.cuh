// fp32-faithful GEMM on the tensor cores ("3xTF32"), used by the inverse-Hessian chain.
#pragma once
#include <cuda_runtime.h>

namespace qt {
namespace tgemm {

// A row-major fp32 matrix given as its tf32 split x = hi + lo (hi tf32-exact; qt_split_tf32 family).
// rows/cols are the extents used for TMA zero-fill (loads) and clipping (stores).
struct Mat {
    const float* hi;
    const float* lo;
    int rows, cols, ld;
};

struct Problem {
    // C[c_row0 + m][c_col0 + n] (op)= sign * sum_k A[a_row0 + m][a_col0 + k] * B[b_row0 + n][b_col0 + k]
    // for m < M, n < N, k < Kd: both operands K-major (contraction over their columns).
    Mat A, B;
    float* C;
    int c_rows, c_cols, ldc;       // extents of the C array (stores are clipped to c_row0+M / c_col0+N)
    int M, N, Kd;                  // Kd % 32 == 0
    int a_row0 = 0, a_col0 = 0, b_row0 = 0, b_col0 = 0, c_row0 = 0, c_col0 = 0;
    int batch = 1;                 // batch b adds b * (stride_r, stride_c) to every offset (M % 128 == 0 and
    int a_sr = 0, a_sc = 0, b_sr = 0, b_sc = 0, c_sr = 0, c_sc = 0;   // N % 256 == 0 when batch > 1)
    bool negate = false;           // sign = -1
    bool accumulate = false;       // C += (TMA reduce-add in L2) instead of C =
    bool lower_tiles_only = false; // skip 128x256 tiles entirely above the diagonal of C (SYRK)
    int a_tri = 0;                 // 1: A[m][k] = 0 for k > m (lower), 2: = 0 for k < m (upper) - k range is trimmed
    int b_tri = 0;                 // same for B[n][k]
    int max_chain = 128;           // longest k span accumulated in TMEM before the partial sum is flushed to C
    int reserve_sms = 0;           // leave this many SMs to other streams (a persistent 148-CTA launch with ~200 KB of
                                   // shared memory per CTA otherwise blocks every 1-10 CTA kernel until it drains)
    // single-product mode: A.hi * B.hi only (plain tf32 GEMM, a third of the tensor work; lo pointers unused)
    bool single = false;
    // bf16 operands (implies single): A.hi / B.hi point at __nv_bfloat16 arrays, ld / cols in elements; kind::f16,
    // fp32 accumulation.  Kd % 64 == 0.
    bool bf16 = false;
    // reduction epilogue (AWQ Gram-form loss): nothing is stored; *loss += sum_{m,n} acc[m][n] * E[c_row0+m][c_col0+n]
    // (E row-major fp32, leading dimension lde; fp64 atomic per warp).  C is ignored.
    const float* E = nullptr;
    int lde = 0;
    double* loss = nullptr;
};

// stream-ordered; returns QT_OK / negative error code
int launch(const Problem& p, cudaStream_t st);

}  // namespace tgemm
}  // namespace qt
