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


// One CTA: Cholesky of an nb x nb diagonal block and its triangular inverse, both blocked in
// shared memory (the block is padded to 128 x 128 with an identity so every loop is uniform):
//   factor : 4 panels of 32 columns, left-looking.  (1) panel -= L[:, :c0] * L[c0:c0+32, :c0]^T with
//            all 256 threads (3x4 register tiles), (2) 32x32 diagonal factor by one warp in registers
//            (right-looking, shuffle broadcasts), (3) rows below solved against it, one thread per row.
//            Measured with clock64: the kernel is bound by the 128 sequential pivots (~450 cycles
//            each for (2), ~350 for (3)), not by throughput.
//   invert : 32x32 diagonal inverses by 4 warps (forward substitution, one column per lane), then
//            two merge levels  Y21 = -Y22 * (L21 * Y11)  as register-tiled shared-memory GEMMs.
// Writes L into A's lower triangle and L^-1 into X's diagonal block.  A non-positive / NaN pivot
// records info = 1-based global column and substitutes 1 (caller applies upstream's Hinv = I).
__global__ void __launch_bounds__(256) potrf_inv_kernel(float* __restrict__ A, float* __restrict__ X, int ld,
                                                        int nb, int kofs, int* __restrict__ info) {
    extern __shared__ float sm[];
    float(*L)[LDS_] = reinterpret_cast<float(*)[LDS_]>(sm);
    float(*Y)[LDS_] = reinterpret_cast<float(*)[LDS_]>(sm + NB * LDS_);
    float(*T)[65] = reinterpret_cast<float(*)[65]>(sm + 2 * NB * LDS_);   // 64 x 64 scratch
    float* dinv = sm + 2 * NB * LDS_ + 64 * 65;                            // 1 / L[j][j]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // batched 128-bit loads: 16 float4 per thread in flight at once (one DRAM/L2 latency, not 64)
    {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int q = tid + 256 * u, i = q >> 5, j4 = (q & 31) * 4;
            v[u] = (i < nb && j4 < nb) ? *reinterpret_cast<const float4*>(A + (long long)(kofs + i) * ld + kofs + j4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const int q = tid + 256 * u, i = q >> 5, j4 = (q & 31) * 4;
            const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int j = j4 + t;
                float x = 0.f;
                if (i < nb && j <= i) x = e[t];
                else if (i >= nb && i == j) x = 1.f;
                L[i][j] = x;
                Y[i][j] = 0.f;
            }
        }
    }
    __syncthreads();
    // ---------------- factor ----------------
    for (int c0 = 0; c0 < NB; c0 += 32) {
        if (c0 > 0) {
            // (1) rows c0..127 (3 per thread at most), columns c0+4tx..+3
            const int tx = tid & 7, ty = tid >> 3;
            float acc[3][4];
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
            const int nrow = (NB - c0) >> 5;   // 3, 2, 1
            for (int k = 0; k < c0; k++) {
                float bv[4];
#pragma unroll
                for (int b = 0; b < 4; b++) bv[b] = L[c0 + 4 * tx + b][k];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (a < nrow) {
                        const float av = L[c0 + ty + 32 * a][k];
#pragma unroll
                        for (int b = 0; b < 4; b++) acc[a][b] = fmaf(av, bv[b], acc[a][b]);
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (a < nrow) {
                    const int r = c0 + ty + 32 * a;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int c = c0 + 4 * tx + b;
                        if (c <= r) L[r][c] -= acc[a][b];
                    }
                }
            }
            __syncthreads();
        }
        // (2) 32 x 32 diagonal block by warp 0, left-looking (Crout): lane = row; column j is
        //     L[i][j] = (a[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j].  Real loops (small code: the
        //     fully unrolled register version stalled on instruction fetch), no store->load
        //     dependence inside the k loop, row reads are conflict-free (stride 129).
        if (warp == 0) {
            // right-looking, the 32 x 32 block lives in registers (lane = row); ~1.1k instructions, the only
            // fully unrolled region of this kernel so it stays inside the instruction cache
            float a[32];
            int badcol = -1;
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = L[c0 + lane][c0 + j];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float d = __shfl_sync(0xffffffffu, a[j], j);
                if (!(d > 0.f)) { d = 1.f; if (badcol < 0) badcol = c0 + j; }
                const float inv = rsqrtf(d);      // ~2 ulp: well inside fp32 Cholesky noise
                const float lij = (lane > j) ? a[j] * inv : 0.f;
                a[j] = (lane == j) ? d * inv : ((lane > j) ? lij : a[j]);
                if (lane == j) dinv[c0 + j] = inv;
#pragma unroll
                for (int k = 0; k < 32; k++) {
                    if (k > j) {
                        const float lkj = __shfl_sync(0xffffffffu, lij, k);
                        a[k] = (lane >= k) ? fmaf(-lij, lkj, a[k]) : a[k];
                    }
                }
            }
            if (badcol >= 0 && lane == 0) atomicCAS(info, 0, kofs + badcol + 1);
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (j <= lane) L[c0 + lane][c0 + j] = a[j];
        }
        __syncthreads();
        // (3) rows below the diagonal block: x * L11^T = a, one thread per row, in place in smem
        if (tid < NB - c0 - 32) {
            float* x = &L[c0 + 32 + tid][c0];
            for (int j = 0; j < 32; j++) {
                const float* lj = &L[c0 + j][c0];
                float s0 = x[j], s1 = 0.f, s2 = 0.f, s3 = 0.f;
                int k = 0;
                for (; k + 3 < j; k += 4) {
                    const float a0 = x[k], a1 = x[k + 1], a2 = x[k + 2], a3 = x[k + 3];
                    const float b0 = lj[k], b1 = lj[k + 1], b2 = lj[k + 2], b3 = lj[k + 3];
                    s0 = fmaf(-a0, b0, s0); s1 = fmaf(-a1, b1, s1);
                    s2 = fmaf(-a2, b2, s2); s3 = fmaf(-a3, b3, s3);
                }
                for (; k < j; k++) s0 = fmaf(-x[k], lj[k], s0);
                x[j] = ((s0 + s1) + (s2 + s3)) * dinv[c0 + j];
            }
        }
        __syncthreads();
    }

    // ---------------- invert ----------------
    // (a) 32 x 32 diagonal inverses: warp w < 4 owns block w, lane c owns column c of the inverse
    //     (forward substitution; Y column reads/writes are lane-consecutive, L reads are broadcasts)
    if (warp < 4) {
        const int b0 = 32 * warp, c = lane;
        for (int i = 0; i < 32; i++) {
            const float* li = &L[b0 + i][b0];
            float s0 = (i == c) ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int k = 0;
            for (; k + 3 < i; k += 4) {
                const float a0 = li[k], a1 = li[k + 1], a2 = li[k + 2], a3 = li[k + 3];
                const float b0_ = Y[b0 + k][b0 + c], b1 = Y[b0 + k + 1][b0 + c], b2 = Y[b0 + k + 2][b0 + c],
                            b3 = Y[b0 + k + 3][b0 + c];
                s0 = fmaf(-a0, b0_, s0); s1 = fmaf(-a1, b1, s1);
                s2 = fmaf(-a2, b2, s2); s3 = fmaf(-a3, b3, s3);
            }
            for (; k < i; k++) s0 = fmaf(-li[k], Y[b0 + k][b0 + c], s0);
            Y[b0 + i][b0 + c] = (i >= c) ? ((s0 + s1) + (s2 + s3)) * dinv[b0 + i] : 0.f;
        }
    }
    __syncthreads();
    // (b) level 32: pairs (0,1) and (2,3); 128 threads per pair, 2 x 4 outputs per thread
    {
        const int pr = tid >> 7, t = tid & 127;
        const int a0 = 64 * pr, b0 = a0 + 32;
        const int tx = t & 7, ty = t >> 3;      // cols 4tx.., rows 2ty..
        float acc[2][4] = {};
        for (int k = 0; k < 32; k++) {          // T = L21 * Y11
            float bv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) bv[q] = Y[a0 + k][a0 + 4 * tx + q];
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const float av = L[b0 + 2 * ty + a][a0 + k];
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][q] = fmaf(av, bv[q], acc[a][q]);
            }
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) T[32 * pr + 2 * ty + a][4 * tx + q] = acc[a][q];
        __syncthreads();
        float acc2[2][4] = {};
        for (int k = 0; k < 32; k++) {          // Y21 = -Y22 * T
            float bv[4];
#pragma unroll
            for (int q = 0; q < 4; q++) bv[q] = T[32 * pr + k][4 * tx + q];
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const float av = Y[b0 + 2 * ty + a][b0 + k];
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[a][q] = fmaf(av, bv[q], acc2[a][q]);
            }
        }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) Y[b0 + 2 * ty + a][a0 + 4 * tx + q] = -acc2[a][q];
    }
    __syncthreads();
    // (c) level 64: one pair of 64 x 64 blocks, 4 x 4 outputs per thread
    {
        const int tx = tid & 15, ty = tid >> 4;
        float acc[4][4] = {};
        for (int k = 0; k < 64; k++) {          // T = L21 * Y11
            float bv[4], av[4];
#pragma unroll
            for (int q = 0; q < 4; q++) bv[q] = Y[k][4 * tx + q];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = L[64 + 4 * ty + a][k];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[a][q] = fmaf(av[a], bv[q], acc[a][q]);
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) T[4 * ty + a][4 * tx + q] = acc[a][q];
        __syncthreads();
        float acc2[4][4] = {};
        for (int k = 0; k < 64; k++) {          // Y21 = -Y22 * T
            float bv[4], av[4];
#pragma unroll
            for (int q = 0; q < 4; q++) bv[q] = T[k][4 * tx + q];
#pragma unroll
            for (int a = 0; a < 4; a++) av[a] = Y[64 + 4 * ty + a][64 + k];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[a][q] = fmaf(av[a], bv[q], acc2[a][q]);
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) Y[64 + 4 * ty + a][4 * tx + q] = -acc2[a][q];
    }
    __syncthreads();
    // write L (zeros above the diagonal: A's upper part is never read) and L^-1, 128-bit stores
#pragma unroll 4
    for (int u = 0; u < 16; u++) {
        const int q = tid + 256 * u, i = q >> 5, j4 = (q & 31) * 4;
        if (i < nb && j4 < nb) {
            const long long g = (long long)(kofs + i) * ld + kofs + j4;
            *reinterpret_cast<float4*>(A + g) = make_float4(L[i][j4], L[i][j4 + 1], L[i][j4 + 2], L[i][j4 + 3]);
            *reinterpret_cast<float4*>(X + g) = make_float4(Y[i][j4], Y[i][j4 + 1], Y[i][j4 + 2], Y[i][j4 + 3]);
        }
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

// Look-ahead resources: the diagonal-block kernel is a single latency-bound CTA (128 sequential
// pivots), so it is issued on a second, high-priority stream as soon as the NEXT panel column has
// received its trailing update, and runs underneath the rest of that update.
struct LookAhead {
    cudaStream_t side = nullptr;
    cudaEvent_t panel_ready = nullptr, potrf_done = nullptr;
    bool ok = false;
};
// one side stream per caller stream: chains that the host runs concurrently on different streams
// must not serialise behind each other's look-ahead kernels
static LookAhead& lookahead(cudaStream_t st) {
    static thread_local struct { cudaStream_t key; LookAhead la; } table[16];
    static thread_local int used = 0;
    for (int i = 0; i < used; i++)
        if (table[i].key == st) return table[i].la;
    const int slot = used < 16 ? used++ : 15;
    LookAhead& la = table[slot].la;
    table[slot].key = st;
    if (!la.side) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        la.ok = cudaStreamCreateWithPriority(&la.side, cudaStreamNonBlocking, hi) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.panel_ready, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.potrf_done, cudaEventDisableTiming) == cudaSuccess;
    }
    return la;
}

static int cholesky_lower(float* A, float* X, int K, int ld, int* info, cudaStream_t st) {
    const size_t smem = (2 * NB * LDS_ + 64 * 65 + NB) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error("potrf smem attr", e); return QT_ERR_CUDA; }
        attr_set = true;
    }
    LookAhead& la = lookahead(st);
    bool potrf_ahead = false;   // potrf of the current block was already issued on the side stream
    for (int k = 0; k < K; k += NB) {
        const int nb = (K - k) < NB ? (K - k) : NB;
        if (potrf_ahead) {
            if (cudaStreamWaitEvent(st, la.potrf_done, 0) != cudaSuccess) return QT_ERR_CUDA;
        } else {
            potrf_inv_kernel<<<1, 256, smem, st>>>(A, X, ld, nb, k, info);
            int rc = check_launch("potrf_inv");
            if (rc) return rc;
        }
        potrf_ahead = false;
        const int rem = K - k - nb;
        if (rem <= 0) break;
        float* P = A + (long long)(k + nb) * ld + k;
        GemmArgs t{};  // TRSM as GEMM: P <- P * (L_kk^-1)^T, in place (single column tile)
        t.A = P; t.B = X + (long long)k * ld + k; t.C = P;
        t.M = rem; t.N = nb; t.Kd = nb; t.lda = t.ldb = t.ldc = ld;
        t.alpha = 1.f; t.beta = 0.f;
        int rc = sgemm(true, t, 1, st);
        if (rc) return rc;
        // SYRK: A22 -= P P^T (lower tiles only), split into the next panel column and the rest
        const int nb2 = rem < NB ? rem : NB;
        GemmArgs s1{};
        s1.A = P; s1.B = P; s1.C = A + (long long)(k + nb) * ld + (k + nb);
        s1.M = rem; s1.N = nb2; s1.Kd = nb; s1.lda = s1.ldb = s1.ldc = ld;
        s1.alpha = -1.f; s1.beta = 1.f; s1.lower_tiles_only = 1;
        rc = sgemm(true, s1, 1, st);
        if (rc) return rc;
        const int rem2 = rem - nb2;
        if (la.ok && rem2 > 0) {
            // next diagonal block is final: factor it on the side stream while the rest updates
            if (cudaEventRecord(la.panel_ready, st) != cudaSuccess) return QT_ERR_CUDA;
            if (cudaStreamWaitEvent(la.side, la.panel_ready, 0) != cudaSuccess) return QT_ERR_CUDA;
            potrf_inv_kernel<<<1, 256, smem, la.side>>>(A, X, ld, nb2, k + nb, info);
            rc = check_launch("potrf_inv(look-ahead)");
            if (rc) return rc;
            if (cudaEventRecord(la.potrf_done, la.side) != cudaSuccess) return QT_ERR_CUDA;
            potrf_ahead = true;
        }
        if (rem2 > 0) {
            float* P2 = P + (long long)nb2 * ld;
            GemmArgs s2{};
            s2.A = P2; s2.B = P2; s2.C = A + (long long)(k + nb + nb2) * ld + (k + nb + nb2);
            s2.M = rem2; s2.N = rem2; s2.Kd = nb; s2.lda = s2.ldb = s2.ldc = ld;
            s2.alpha = -1.f; s2.beta = 1.f; s2.lower_tiles_only = 1;
            rc = sgemm(true, s2, 1, st);
            if (rc) return rc;
        }
    }
    return QT_OK;
}

// X holds the inverses of the NB x NB diagonal blocks of L.  Merge pairs of blocks level by
// level:  inv([[A,0],[C,B]]) = [[A^-1,0],[-B^-1 C A^-1, B^-1]].  All pairs of one level run as
// one batched GEMM (+ one more launch for a ragged last pair).
static int trtri_lower(const float* L, float* X, float* W, int K, int ld, cudaStream_t st) {
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
            const long long stride = 2 * s * ((long long)ld + 1);
            GemmArgs g1{};  // T = C * A^-1   (A^-1 lower-triangular as the [k][n] operand)
            g1.A = L + b0 * ld + a0; g1.B = X + a0 * ld + a0; g1.C = W + b0 * ld + a0;
            g1.M = (int)sB; g1.N = (int)s; g1.Kd = (int)s; g1.lda = g1.ldb = g1.ldc = ld;
            g1.alpha = 1.f; g1.beta = 0.f; g1.b_lower_tri = 1;
            g1.strideA = g1.strideB = g1.strideC = stride;
            int rc = sgemm(false, g1, (int)batch, st);
            if (rc) return rc;
            GemmArgs g2{};  // X[C] = -B^-1 * T   (B^-1 lower-triangular as the [m][k] operand)
            g2.A = X + b0 * ld + b0; g2.B = W + b0 * ld + a0; g2.C = X + b0 * ld + a0;
            g2.M = (int)sB; g2.N = (int)s; g2.Kd = (int)sB; g2.lda = g2.ldb = g2.ldc = ld;
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

int qt_sgemm_ex(int b_is_nk, const float* A, const float* B, float* C, int M, int N, int Kd, int lda, int ldb, int ldc,
                float alpha, float beta, int lower_tiles_only, int a_lower_tri, int b_lower_tri, int tri_row_offset,
                void* stream);

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
    int rc = cholesky_lower(A, X, K, K, info, st);
    if (rc) return rc;
    rc = trtri_lower(A, X, W, K, K, st);
    if (rc) return rc;
    dim3 grid((K + 255) / 256, K);
    flip_upper_kernel<<<grid, 256, 0, st>>>(X, A, K);
    return check_launch("flip_upper");
}

// Building blocks of the multi-GPU chain (engine/pipeline.py): Cholesky + triangular inverse of an n x n
// block living inside larger buffers of leading dimension ld (A: in lower(H block) -> out L; X: out L^-1;
// W: scratch), and the final index reversal.  info is accumulated (first failure wins), not reset.
int qt_tri_chain_block(float* A, float* X, float* W, int n, int ld, int* info, void* stream) {
    if (!A || !X || !W || !info || n <= 0 || (n & 3) || (ld & 3) || ld < n) return QT_ERR_INVALID;
    if (((uintptr_t)A & 15) || ((uintptr_t)X & 15) || ((uintptr_t)W & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = cholesky_lower(A, X, n, ld, info, st);
    if (rc) return rc;
    return trtri_lower(A, X, W, n, ld, st);
}

int qt_flip_upper(const float* X, float* U, int K, void* stream) {
    if (!X || !U || K <= 0) return QT_ERR_INVALID;
    dim3 grid((K + 255) / 256, K);
    flip_upper_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, U, K);
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
    return qt_sgemm_ex(b_is_nk, A, B, C, M, N, Kd, lda, ldb, ldc, alpha, beta, lower_tiles_only, a_lower_tri, b_lower_tri,
                       0, stream);
}

// tri_row_offset: the M rows of this call start at that row of the triangular structure (row-sliced SYRK /
// triangular-A products on one rank of a multi-GPU split)
int qt_sgemm_ex(int b_is_nk, const float* A, const float* B, float* C, int M, int N, int Kd, int lda, int ldb, int ldc,
                float alpha, float beta, int lower_tiles_only, int a_lower_tri, int b_lower_tri, int tri_row_offset,
                void* stream) {
    GemmArgs g{};
    g.tri_row_offset = tri_row_offset;
    g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.Kd = Kd; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.alpha = alpha; g.beta = beta;
    g.lower_tiles_only = lower_tiles_only; g.a_lower_tri = a_lower_tri; g.b_lower_tri = b_lower_tri;
    return sgemm(b_is_nk != 0, g, 1, (cudaStream_t)stream);
}

}  // extern "C"
