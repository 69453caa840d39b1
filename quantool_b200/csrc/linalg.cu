// fp32 inverse-Hessian factor for GPTQ on sm_100a.
//
// Replaces UPSTREAM llmcompressor gptq_quantize.py `quantize_weight` prologue (SURVEY.md §A.3,
// row a2), reached from ref/src/quantool/methods/llm_compressor/gptq/gptq.py:86 via
// llmcompressor.oneshot at ref/src/quantool/methods/llm_compressor/base.py:159-161:
//     H += damp*I ; L = cholesky(H) ; Hinv = cholesky_inverse(L) ; U = cholesky(Hinv, upper)
//
// Same result, half the flops: U = R^-1 where H = R R^T with R upper ("reverse" Cholesky).
// With J the index reversal, J H J = Lf Lf^T (ordinary lower Cholesky) and
// U[i][j] = (Lf^-1)[K-1-i][K-1-j].  So: gather H flipped (fused with the act_order permutation
// and damping), one blocked right-looking Cholesky, one recursive blocked triangular inverse,
// one flip.  All trailing updates are fp32 FFMA GEMMs (sgemm.cuh); nothing leaves fp32.
//
// HBM layout: every matrix is K x K fp32 row-major, leading dimension K.  Three buffers:
//   A  in: flipped damped H (lower triangle read)   out: U (upper triangle, zeros below)
//   X  scratch: Lf^-1 (lower)                       W  scratch: GEMM temporaries
#include "sgemm.cuh"

namespace qt {
namespace linalg {

constexpr int NB = 128;
constexpr int LDS_ = NB + 1;

// One CTA: Cholesky of an nb x nb diagonal block (right-looking, in shared memory), then its
// triangular inverse.  Writes L back into A's lower triangle and L^-1 into X's diagonal block.
// A non-positive / NaN pivot records info = 1-based global column and substitutes 1 so that the
// caller can fall back the way upstream does on LinAlgError (Hinv = I).
__global__ void __launch_bounds__(256) potrf_inv_kernel(float* __restrict__ A, float* __restrict__ X, int ld,
                                                        int nb, int kofs, int* __restrict__ info) {
    extern __shared__ float sm[];
    float(*L)[LDS_] = reinterpret_cast<float(*)[LDS_]>(sm);
    float(*Y)[LDS_] = reinterpret_cast<float(*)[LDS_]>(sm + NB * LDS_);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < nb * nb; idx += 256) {
        const int i = idx / nb, j = idx - i * nb;
        L[i][j] = (j <= i) ? A[(long long)(kofs + i) * ld + kofs + j] : 0.f;
        Y[i][j] = 0.f;
    }
    __syncthreads();
    for (int j = 0; j < nb; j++) {
        if (tid == 0) {
            float d = L[j][j];
            if (!(d > 0.f)) {
                atomicCAS(info, 0, kofs + j + 1);
                d = 1.f;
            }
            L[j][j] = sqrtf(d);
        }
        __syncthreads();
        const float d = L[j][j];
        for (int i = j + 1 + tid; i < nb; i += 256) L[i][j] = L[i][j] / d;
        __syncthreads();
        const int rem = nb - j - 1;
        // rows j+1..nb-1; two threads per row, each walking half of the row's lower part
        for (int r = tid >> 1; r < rem; r += 128) {
            const int i = j + 1 + r;
            const float lij = L[i][j];
            for (int k = j + 1 + (tid & 1); k <= i; k += 2) L[i][k] = fmaf(-lij, L[k][j], L[i][k]);
        }
        __syncthreads();
    }
    // triangular inverse: thread c owns column c of Y = L^-1 (forward substitution)
    if (tid < nb) {
        const int c = tid;
        for (int i = c; i < nb; i++) {
            float s0 = (i == c) ? 1.f : 0.f, s1 = 0.f;
            int k = c;
            for (; k + 1 < i; k += 2) {
                s0 = fmaf(-L[i][k], Y[k][c], s0);
                s1 = fmaf(-L[i][k + 1], Y[k + 1][c], s1);
            }
            if (k < i) s0 = fmaf(-L[i][k], Y[k][c], s0);
            Y[i][c] = (s0 + s1) / L[i][i];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < nb * nb; idx += 256) {
        const int i = idx / nb, j = idx - i * nb;
        if (j <= i) A[(long long)(kofs + i) * ld + kofs + j] = L[i][j];
        X[(long long)(kofs + i) * ld + kofs + j] = (j <= i) ? Y[i][j] : 0.f;
    }
}

// U[i][j] = X[K-1-i][K-1-j] for j >= i, 0 below the diagonal
__global__ void __launch_bounds__(256) flip_upper_kernel(const float* __restrict__ X, float* __restrict__ U, int K) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= K) return;
    U[(long long)i * K + j] = (j >= i) ? X[(long long)(K - 1 - i) * K + (K - 1 - j)] : 0.f;
}

// diag statistics: dead[i] = (H[i][i] == 0); damp = percdamp * mean(diag with dead -> 1)
__global__ void __launch_bounds__(1024) diag_stats_kernel(const float* __restrict__ H, int K, float percdamp,
                                                          uint8_t* __restrict__ dead, float* __restrict__ damp_out) {
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < K; i += 1024) {
        float d = H[(long long)i * K + i];
        const bool dd = (d == 0.f);
        dead[i] = dd ? 1 : 0;
        if (dd) d = 1.f;
        s += d;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = red[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) *damp_out = percdamp * (v / (float)K);
    }
}

// Hf[i][j] = H'[p(i)][p(j)] + damp*(i==j), p(i) = perm[K-1-i] (perm == nullptr: identity),
// H' = H with dead diagonal entries set to 1.  Only j <= i is written (lower triangle).
__global__ void __launch_bounds__(256) gather_flip_kernel(const float* __restrict__ H, const int* __restrict__ perm,
                                                          const uint8_t* __restrict__ dead,
                                                          const float* __restrict__ damp, float* __restrict__ Hf, int K) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= K || j > i) return;
    const int pi = perm ? perm[K - 1 - i] : (K - 1 - i);
    const int pj = perm ? perm[K - 1 - j] : (K - 1 - j);
    float v = H[(long long)pi * K + pj];
    if (i == j) {
        if (dead[pi]) v = 1.f;
        v += *damp;
    }
    Hf[(long long)i * K + j] = v;
}

__global__ void __launch_bounds__(256) set_identity_kernel(float* __restrict__ U, int K) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = blockIdx.y;
    if (j < K) U[(long long)i * K + j] = (i == j) ? 1.f : 0.f;
}

static int cholesky_lower(float* A, float* X, int K, int* info, cudaStream_t st) {
    const size_t smem = 2 * NB * LDS_ * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error("potrf smem attr", e); return QT_ERR_CUDA; }
        attr_set = true;
    }
    for (int k = 0; k < K; k += NB) {
        const int nb = (K - k) < NB ? (K - k) : NB;
        potrf_inv_kernel<<<1, 256, smem, st>>>(A, X, K, nb, k, info);
        int rc = check_launch("potrf_inv");
        if (rc) return rc;
        const int rem = K - k - nb;
        if (rem <= 0) break;
        float* P = A + (long long)(k + nb) * K + k;
        GemmArgs t{};  // TRSM as GEMM: P <- P * (L_kk^-1)^T, in place (single column tile)
        t.A = P; t.B = X + (long long)k * K + k; t.C = P;
        t.M = rem; t.N = nb; t.Kd = nb; t.lda = t.ldb = t.ldc = K;
        t.alpha = 1.f; t.beta = 0.f;
        rc = sgemm(true, t, 1, st);
        if (rc) return rc;
        GemmArgs s{};  // SYRK: A22 -= P P^T, lower tiles only
        s.A = P; s.B = P; s.C = A + (long long)(k + nb) * K + (k + nb);
        s.M = rem; s.N = rem; s.Kd = nb; s.lda = s.ldb = s.ldc = K;
        s.alpha = -1.f; s.beta = 1.f; s.lower_tiles_only = 1;
        rc = sgemm(true, s, 1, st);
        if (rc) return rc;
    }
    return QT_OK;
}

// X holds the inverses of the NB x NB diagonal blocks of L.  Merge pairs of blocks level by
// level:  inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]].  All pairs of one level run as
// one batched GEMM (+ one more launch for a ragged last pair).
static int trtri_lower(const float* L, float* X, float* W, int K, cudaStream_t st) {
    for (long long s = NB; s < K; s *= 2) {
        const long long nblk = (K + s - 1) / s;
        const long long npairs = nblk / 2;  // pairs whose B block exists
        if (npairs == 0) break;
        // last pair is ragged if its B block is shorter than s
        const long long lastB0 = (2 * (npairs - 1) + 1) * s;
        const long long lastSB = (K - lastB0) < s ? (K - lastB0) : s;
        const long long nfull = (lastSB == s) ? npairs : npairs - 1;
        for (int pass = 0; pass < 2; pass++) {
            const long long batch = pass == 0 ? nfull : (npairs - nfull);
            if (batch <= 0) continue;
            const long long p0 = pass == 0 ? 0 : nfull;
            const long long sB = pass == 0 ? s : lastSB;
            const long long a0 = 2 * p0 * s, b0 = a0 + s;
            const long long stride = 2 * s * ((long long)K + 1);
            GemmArgs g1{};  // T = C * A^-1   (A^-1 lower-triangular as the [k][n] operand)
            g1.A = L + b0 * K + a0; g1.B = X + a0 * K + a0; g1.C = W + b0 * K + a0;
            g1.M = (int)sB; g1.N = (int)s; g1.Kd = (int)s; g1.lda = g1.ldb = g1.ldc = K;
            g1.alpha = 1.f; g1.beta = 0.f; g1.b_lower_tri = 1;
            g1.strideA = g1.strideB = g1.strideC = stride;
            int rc = sgemm(false, g1, (int)batch, st);
            if (rc) return rc;
            GemmArgs g2{};  // X[C] = -B^-1 * T   (B^-1 lower-triangular as the [m][k] operand)
            g2.A = X + b0 * K + b0; g2.B = W + b0 * K + a0; g2.C = X + b0 * K + a0;
            g2.M = (int)sB; g2.N = (int)s; g2.Kd = (int)sB; g2.lda = g2.ldb = g2.ldc = K;
            g2.alpha = -1.f; g2.beta = 0.f; g2.a_lower_tri = 1;
            g2.strideA = g2.strideB = g2.strideC = stride;
            rc = sgemm(false, g2, (int)batch, st);
            if (rc) return rc;
        }
    }
    return QT_OK;
}

}  // namespace linalg
}  // namespace qt

using namespace qt;
using namespace qt::linalg;

extern "C" {

int qt_gptq_prepare_hessian(const float* H, const int* perm, int K, float percdamp, float* Hf, uint8_t* dead,
                            float* damp_scratch, void* stream) {
    if (!H || !Hf || !dead || !damp_scratch || K <= 0 || (K & 3)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    diag_stats_kernel<<<1, 1024, 0, st>>>(H, K, percdamp, dead, damp_scratch);
    int rc = check_launch("diag_stats");
    if (rc) return rc;
    dim3 grid((K + 255) / 256, K);
    gather_flip_kernel<<<grid, 256, 0, st>>>(H, perm, dead, damp_scratch, Hf, K);
    return check_launch("gather_flip");
}

int qt_gptq_hinv_factor(float* A, float* X, float* W, int K, int* info, void* stream) {
    if (!A || !X || !W || !info || K <= 0 || (K & 3)) return QT_ERR_INVALID;
    if (((uintptr_t)A & 15) || ((uintptr_t)X & 15) || ((uintptr_t)W & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) { set_last_error("memset info", e); return QT_ERR_CUDA; }
    int rc = cholesky_lower(A, X, K, info, st);
    if (rc) return rc;
    rc = trtri_lower(A, X, W, K, st);
    if (rc) return rc;
    dim3 grid((K + 255) / 256, K);
    flip_upper_kernel<<<grid, 256, 0, st>>>(X, A, K);
    return check_launch("flip_upper");
}

int qt_set_identity(float* U, int K, void* stream) {
    if (!U || K <= 0) return QT_ERR_INVALID;
    dim3 grid((K + 255) / 256, K);
    set_identity_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(U, K);
    return check_launch("set_identity");
}

// exposed for tests and for the bench's roofline probe of the fp32 GEMM
int qt_sgemm(int b_is_nk, const float* A, const float* B, float* C, int M, int N, int Kd, int lda, int ldb, int ldc,
             float alpha, float beta, int lower_tiles_only, int a_lower_tri, int b_lower_tri, void* stream) {
    GemmArgs g{};
    g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.Kd = Kd; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta;
    g.lower_tiles_only = lower_tiles_only; g.a_lower_tri = a_lower_tri; g.b_lower_tri = b_lower_tri;
    return sgemm(b_is_nk != 0, g, 1, (cudaStream_t)stream);
}

}  // extern "C"
