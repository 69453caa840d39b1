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
// one flip.  Two variants of the GEMM work: fp32 FFMA GEMMs (sgemm.cuh) everywhere, or - when K is a multiple
// of 256 and the caller provides four more K x K workspaces - the big products on the tensor cores as
// fp32-faithful 3xTF32 GEMMs (tgemm.cu): the Cholesky trailing update is deferred per 512-column outer block
// (inner dimension 512 instead of 128: 4x less C traffic) and the block merges of the triangular inverse run as
// S = X11^T-form products with triangular k trimming.  Panels, TRSMs and the 128-level merges stay FFMA.
//
// HBM layout: every matrix is K x K fp32 row-major, leading dimension K.  Three buffers:
//   A  in: flipped damped H (lower triangle read)   out: U (upper triangle, zeros below)
//   X  scratch: Lf^-1 (lower)                       W  scratch: GEMM temporaries
#include "sgemm.cuh"
#include "tgemm.cuh"

namespace qt {
namespace linalg {

constexpr int NB = 128;
constexpr int LDS_ = NB + 1;


// One CTA: Cholesky of an nb x nb diagonal block and its triangular inverse, both blocked in
// shared memory (the block is padded to 128 x 128 with an identity so every loop is uniform):
//   factor : 4 panels of 32 columns, left-looking.  (1) panel -= L[:, :c0] * L[c0:c0+32, :c0]^T with
//            all 256 threads (3x4 register tiles), (2) 32x32 diagonal factor by one warp in registers
//            (right-looking, shuffle broadcasts, no lane masks), (2b) its inverse by the same warp,
//            (3) rows below = (rows) * inverse^T as a product over all threads.  The kernel is bound by the
//            128 sequential pivots; the first version (masked updates: ~450 cycles per pivot, rows below
//            solved one thread per row: ~350 cycles per pivot column) took 86 us.
//   invert : the 32x32 diagonal inverses come from (2b); two merge levels  Y21 = -Y22 * (L21 * Y11)  as
//            register-tiled shared-memory GEMMs.
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
        // (2) 32 x 32 diagonal block by warp 0, right-looking, the block lives in registers (lane = row).  No lane
        //     masks in the rank-1 update: entries above the diagonal (column k in a lane < k) are never read again,
        //     so every lane runs the same shuffle + FMA stream (the masked form cost ~450 cycles per pivot).
        //     (2b) the same warp then inverts the factor by forward substitution (lane = column of the inverse,
        //     kept in registers; L rows are shared-memory broadcasts).  That inverse is the diagonal block of the
        //     final L^-1 AND turns the solve of the rows below into a product (3).
        if (warp == 0) {
            float a[32];
            int badcol = -1;
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = L[c0 + lane][c0 + j];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                float d = __shfl_sync(0xffffffffu, a[j], j);
                if (!(d > 0.f)) { d = 1.f; if (badcol < 0) badcol = c0 + j; }
                const float inv = rsqrtf(d);      // ~2 ulp: well inside fp32 Cholesky noise
                const float lij = a[j] * inv;
                a[j] = (lane == j) ? d * inv : lij;
                if (lane == j) dinv[c0 + j] = inv;
#pragma unroll
                for (int k = j + 1; k < 32; k++) {
                    const float lkj = __shfl_sync(0xffffffffu, lij, k);
                    a[k] = fmaf(-lij, lkj, a[k]);
                }
            }
            if (badcol >= 0 && lane == 0) atomicCAS(info, 0, kofs + badcol + 1);
#pragma unroll
            for (int j = 0; j < 32; j++)
                if (j <= lane) L[c0 + lane][c0 + j] = a[j];
            __syncwarp();
            // (2b) Y11 = L11^-1, column `lane`: y[i] = (delta - sum_{k<i} L[i][k] y[k]) / L[i][i]
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const float* li = &L[c0 + i][c0];
                float s0 = (i == lane) ? 1.f : 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int k = 0; k < i; k++) {
                    const float t = li[k];
                    if ((k & 3) == 0) s0 = fmaf(-t, y[k], s0);
                    else if ((k & 3) == 1) s1 = fmaf(-t, y[k], s1);
                    else if ((k & 3) == 2) s2 = fmaf(-t, y[k], s2);
                    else s3 = fmaf(-t, y[k], s3);
                }
                y[i] = (i >= lane) ? ((s0 + s1) + (s2 + s3)) * dinv[c0 + i] : 0.f;
                Y[c0 + i][c0 + lane] = y[i];
            }
        }
        __syncthreads();
        // (3) rows below the diagonal block: P = A[:, c0:c0+32] * Y11^T (Y11 lower-triangular: k <= c).  Thread =
        //     (row, column parity): 96 rows x 2 at most; the row segment is read into registers before any thread
        //     of the row writes (barrier in between).
        {
            const int nbelow = NB - c0 - 32;
            const int rrow = tid % 96, par = tid / 96;           // par is warp-uniform (96 = 3 warps)
            const bool act = par < 2 && rrow < nbelow;
            float o[16];
            if (act) {
                float x[32];
                const float* xr = &L[c0 + 32 + rrow][c0];
#pragma unroll
                for (int k = 0; k < 32; k++) x[k] = xr[k];
#pragma unroll
                for (int cc = 0; cc < 16; cc++) {
                    const int c = 2 * cc + par;
                    const float* yc = &Y[c0 + c][c0];
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        if (k <= 2 * cc + 1) {                    // k <= c for par = 1; one harmless zero term for par = 0
                            if (k & 1) s1 = fmaf(x[k], yc[k], s1); else s0 = fmaf(x[k], yc[k], s0);
                        }
                    }
                    o[cc] = s0 + s1;
                }
            }
            __syncthreads();
            if (act) {
                float* xr = &L[c0 + 32 + rrow][c0];
#pragma unroll
                for (int cc = 0; cc < 16; cc++) xr[2 * cc + par] = o[cc];
            }
        }
        __syncthreads();
    }

    // ---------------- invert ----------------
    // the 32 x 32 diagonal inverses are already in Y (2b); two merge levels  Y21 = -Y22 * (L21 * Y11)
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

// The two single-tile products on the blocked Cholesky's critical path, in ONE launch over 10 CTAs:
//     P_top <- P_top * (L_kk^-1)^T          (TRSM of the nb2 rows right below the diagonal block, as a product)
//     D     <- D - P_top * P_top^T          (the next diagonal block, lower 32 x 32 tiles)
// They used to be two 128 x 128 x 128 FFMA GEMM launches of one CTA each (24-35 us apiece, CUPTI) between every
// pair of potrf_inv launches.  CTA (ti, tj), tj <= ti, owns tile (ti, tj) of D and computes the two 32-row strips of
// the solved panel it needs itself (redundantly across CTAs: no grid-wide dependency); the diagonal CTAs write their
// strip back as the new P_top.  fp32 FFMA, fixed summation order.
constexpr int PT_LD = NB + 1;
constexpr size_t PT_SMEM = (size_t)(NB + 4 * 32) * PT_LD * sizeof(float);

__device__ __forceinline__ void panel_strip(const float (*Ps)[PT_LD], const float (*Xs)[PT_LD], float (*Ss)[PT_LD], int tid) {
    // S[r][c] = sum_{j <= c} P[r][j] X[c][j];  thread = rows 4 rg .. +3, columns lane + 32 m
    const int lane = tid & 31, rg = tid >> 5;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int m = 0; m < 4; m++) acc[a][m] = 0.f;
#pragma unroll
    for (int jb = 0; jb < 4; jb++) {
#pragma unroll 4
        for (int jj = 0; jj < 32; jj++) {
            const int j = 32 * jb + jj;
            float pv[4], xv[4];
#pragma unroll
            for (int a = 0; a < 4; a++) pv[a] = Ps[4 * rg + a][j];
#pragma unroll
            for (int m = 0; m < 4; m++)
                if (m >= jb) xv[m] = Xs[lane + 32 * m][j];       // columns c < 32 jb have X[c][j] = 0 for this j
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int m = 0; m < 4; m++)
                    if (m >= jb) acc[a][m] = fmaf(pv[a], xv[m], acc[a][m]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int m = 0; m < 4; m++) Ss[4 * rg + a][lane + 32 * m] = acc[a][m];
}

__global__ void __launch_bounds__(256) panel_top_kernel(float* __restrict__ A, float* __restrict__ X, int ld, int k,
                                                        int nb, int nb2) {
    extern __shared__ float sm[];
    float(*Xs)[PT_LD] = reinterpret_cast<float(*)[PT_LD]>(sm);
    float(*Pi)[PT_LD] = reinterpret_cast<float(*)[PT_LD]>(sm + NB * PT_LD);
    float(*Pj)[PT_LD] = reinterpret_cast<float(*)[PT_LD]>(sm + (NB + 32) * PT_LD);
    float(*Si)[PT_LD] = reinterpret_cast<float(*)[PT_LD]>(sm + (NB + 64) * PT_LD);
    float(*Sj)[PT_LD] = reinterpret_cast<float(*)[PT_LD]>(sm + (NB + 96) * PT_LD);
    const int tid = threadIdx.x;
    int ti = 0, tj = blockIdx.x;                     // lower-triangular tile index -> (ti, tj)
    while (tj > ti) { tj -= ti + 1; ti++; }
    if (32 * ti >= nb2) return;
    const int r0 = k + nb;                           // first row of the panel's top rows (= next diagonal block)
    // X_kk (lower-triangular inverse of the diagonal block), zero-padded
    for (int q = tid; q < NB * (NB / 4); q += 256) {
        const int i = q >> 5, j4 = (q & 31) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nb && j4 < nb) v = *reinterpret_cast<const float4*>(X + (long long)(k + i) * ld + k + j4);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; t++) Xs[i][j4 + t] = (j4 + t <= i && j4 + t < nb) ? e[t] : 0.f;
    }
    for (int q = tid; q < 2 * 32 * (NB / 4); q += 256) {
        const int which = q >> 10, rr = (q >> 5) & 31, j4 = (q & 31) * 4;
        const int row = 32 * (which ? tj : ti) + rr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < nb2 && j4 < nb) v = *reinterpret_cast<const float4*>(A + (long long)(r0 + row) * ld + k + j4);
        float(*P)[PT_LD] = which ? Pj : Pi;
        P[rr][j4] = v.x; P[rr][j4 + 1] = v.y; P[rr][j4 + 2] = v.z; P[rr][j4 + 3] = v.w;
    }
    __syncthreads();
    panel_strip(Pi, Xs, Si, tid);
    if (tj != ti) panel_strip(Pj, Xs, Sj, tid);
    __syncthreads();
    const float(*Sb)[PT_LD] = (tj != ti) ? Sj : Si;
    if (tj == ti) {
        // The solved strip is the new P_top - but other CTAs of this grid may not have read the ORIGINAL rows yet,
        // so it goes to the same position of X (below diagonal block k: unused until the triangular inverse
        // overwrites it) and panel_top_commit_kernel copies it into A afterwards, off the critical path.
        for (int q = tid; q < 32 * (NB / 4); q += 256) {
            const int rr = q >> 5, j4 = (q & 31) * 4;
            const int row = 32 * ti + rr;
            if (row < nb2 && j4 < nb)
                *reinterpret_cast<float4*>(X + (long long)(r0 + row) * ld + k + j4) =
                    make_float4(Si[rr][j4], Si[rr][j4 + 1], Si[rr][j4 + 2], Si[rr][j4 + 3]);
        }
    }
    // D tile (ti, tj) -= Si * Sb^T : thread = row tid / 8, columns tid % 8 + 8 m
    {
        const int r = tid >> 3, cg = tid & 7;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
        for (int j = 0; j < NB; j++) {
            const float av = Si[r][j];
#pragma unroll
            for (int m = 0; m < 4; m++) acc[m] = fmaf(av, Sb[cg + 8 * m][j], acc[m]);
        }
        const int grow = 32 * ti + r;
        if (grow < nb2) {
            float* drow = A + (long long)(r0 + grow) * ld + r0 + 32 * tj;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int c = cg + 8 * m;
                if (32 * tj + c < nb2) drow[c] -= acc[m];
            }
        }
    }
}

__global__ void __launch_bounds__(256) panel_top_commit_kernel(float* __restrict__ A, const float* __restrict__ X, int ld,
                                                               int k, int nb, int nb2, float* __restrict__ Lh,
                                                               float* __restrict__ Ll) {
    const int r0 = k + nb;
    for (int q = blockIdx.x * 256 + threadIdx.x; q < nb2 * (NB / 4); q += gridDim.x * 256) {
        const int row = q >> 5, j4 = (q & 31) * 4;
        if (j4 < nb) {
            const long long g = (long long)(r0 + row) * ld + k + j4;
            const float4 v = *reinterpret_cast<const float4*>(X + g);
            *reinterpret_cast<float4*>(A + g) = v;
            if (Lh) {     // tf32 split of the solved rows (operands of the deferred tensor-core updates)
                float4 h, l;
                tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y);
                tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
                *reinterpret_cast<float4*>(Lh + g) = h;
                *reinterpret_cast<float4*>(Ll + g) = l;
            }
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

// Same result, one CTA per output row: the source row H[p(i)][:] is read once, coalesced, into shared memory and the
// permuted gather happens there (the per-element global gather moved a 32-byte sector per 4-byte element through
// L2: 1.1 ms at K = 14336, on the multi-GPU critical path between the all-reduce of H and the chain).
__global__ void __launch_bounds__(512) gather_flip_row_kernel(const float* __restrict__ H, const int* __restrict__ perm,
                                                              const uint8_t* __restrict__ dead,
                                                              const float* __restrict__ damp, float* __restrict__ Hf, int K) {
    extern __shared__ float srow[];
    const int i = blockIdx.x;
    const int pi = perm ? perm[K - 1 - i] : (K - 1 - i);
    const float4* src = reinterpret_cast<const float4*>(H + (long long)pi * K);
    for (int c = threadIdx.x; c < (K >> 2); c += 512) reinterpret_cast<float4*>(srow)[c] = src[c];
    __syncthreads();
    float* dst = Hf + (long long)i * K;
    for (int j = threadIdx.x; j <= i; j += 512) {
        const int pj = perm ? perm[K - 1 - j] : (K - 1 - j);
        float v = srow[pj];
        if (j == i) {
            if (dead[pi]) v = 1.f;
            v += *damp;
        }
        dst[j] = v;
    }
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
    // tensor-core Cholesky: the far part of a deferred update runs here, underneath the next outer block's panels
    cudaStream_t far = nullptr;
    cudaEvent_t far_ready = nullptr, far_done = nullptr;
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
                cudaEventCreateWithFlags(&la.potrf_done, cudaEventDisableTiming) == cudaSuccess &&
                cudaStreamCreateWithFlags(&la.far, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.far_ready, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.far_done, cudaEventDisableTiming) == cudaSuccess;
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

// ---------------------------------------------------------------------------------------------------
// tensor-core variant
// ---------------------------------------------------------------------------------------------------
// tf32 split of a rows x cols block (optionally transposed), batched over diagonal blocks.
//   TRANSPOSE = false: dst[r][c] = split(src[r][c]);  TRANSPOSE = true: dst[c][r] = split(src[r][c])
//   tri (in the SOURCE block's coordinates): 1 keeps c <= r (lower), 0 keeps everything; the rest is written as 0
template <bool TRANSPOSE>
__global__ void __launch_bounds__(256) split_block_kernel(const float* __restrict__ src, int lds, float* __restrict__ hi,
                                                          float* __restrict__ lo, int ldd, int rows, int cols, int tri,
                                                          long long sstride, long long dstride) {
    __shared__ float t[32][33];
    src += blockIdx.z * sstride;
    hi += blockIdx.z * dstride;
    lo += blockIdx.z * dstride;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!TRANSPOSE) {
        for (int rr = ty; rr < 32; rr += 8) {
            const int r = r0 + rr, c = c0 + tx;
            if (r < rows && c < cols) {
                float h = 0.f, l = 0.f;
                if (!tri || c <= r) tf32_split(src[(long long)r * lds + c], h, l);
                hi[(long long)r * ldd + c] = h;
                lo[(long long)r * ldd + c] = l;
            }
        }
    } else {
        for (int rr = ty; rr < 32; rr += 8) {
            const int r = r0 + rr, c = c0 + tx;
            t[rr][tx] = (r < rows && c < cols && (!tri || c <= r)) ? src[(long long)r * lds + c] : 0.f;
        }
        __syncthreads();
        for (int cc = ty; cc < 32; cc += 8) {
            const int c = c0 + cc, r = r0 + tx;     // destination row c, column r
            if (c < cols && r < rows) {
                float h, l;
                tf32_split(t[tx][cc], h, l);
                hi[(long long)c * ldd + r] = h;
                lo[(long long)c * ldd + r] = l;
            }
        }
    }
}

static int split_block(bool transpose, const float* src, int lds, float* hi, float* lo, int ldd, int rows, int cols,
                       int tri, int batch, long long sstride, long long dstride, cudaStream_t st) {
    if (rows <= 0 || cols <= 0 || batch <= 0) return QT_OK;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    if (transpose) split_block_kernel<true><<<grid, 256, 0, st>>>(src, lds, hi, lo, ldd, rows, cols, tri, sstride, dstride);
    else split_block_kernel<false><<<grid, 256, 0, st>>>(src, lds, hi, lo, ldd, rows, cols, tri, sstride, dstride);
    return check_launch("split_block");
}

constexpr int OB = 512;   // outer block of the tensor-core Cholesky: trailing update deferred over 4 panels

// The tensor cores accumulate in fp32 with truncation, which is harmless for sums of mixed sign but biases a
// long sum of squares: the diagonal of a SYRK update (measured: ~2e-3 absolute on sums of ~500, i.e. ~1000 ulp).
// The diagonal entries of the deferred update are therefore computed here with FFMA in round-to-nearest -
// dfix[i] = A[i][i] - sum_k L[i][k]^2 before the tensor-core GEMM, written back after it.  One warp per row.
__global__ void __launch_bounds__(256) diag_fix_pre_kernel(const float* __restrict__ A, int ld, int row0, int nrows,
                                                           int col0, int ncols, float* __restrict__ dfix) {
    const int r = row0 + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= row0 + nrows) return;
    const int lane = threadIdx.x & 31;
    const float* p = A + (long long)r * ld + col0;
    float s0 = 0.f, s1 = 0.f;
    for (int c = lane * 4; c < ncols; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(p + c);
        s0 = fmaf(v.x, v.x, s0); s1 = fmaf(v.y, v.y, s1);
        s0 = fmaf(v.z, v.z, s0); s1 = fmaf(v.w, v.w, s1);
    }
    const float s = warp_sum(s0 + s1);
    if (lane == 0) dfix[r] = A[(long long)r * ld + r] - s;
}
__global__ void __launch_bounds__(256) diag_fix_post_kernel(float* __restrict__ A, int ld, int row0, int nrows,
                                                            const float* __restrict__ dfix) {
    const int r = row0 + blockIdx.x * 256 + threadIdx.x;
    if (r < row0 + nrows) A[(long long)r * ld + r] = dfix[r];
}

// Blocked Cholesky with the trailing update on the tensor cores.  Lh/Ll receive the tf32 split of every
// sub-diagonal panel of L (what the deferred updates and, later, the triangular inverse read as operands).
static int cholesky_lower_tc(float* A, float* X, float* Lh, float* Ll, float* dfix, int K, int ld, int* info,
                             cudaStream_t st) {
    const size_t smem = (2 * NB * LDS_ + 64 * 65 + NB) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(panel_top_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM);
    if (e != cudaSuccess) { set_last_error("potrf smem attr", e); return QT_ERR_CUDA; }
    LookAhead& la = lookahead(st);
    bool potrf_ahead = false, far_pending = false;
    for (int ob = 0; ob < K; ob += OB) {
        const int oe = (ob + OB) < K ? (ob + OB) : K;
        for (int k = ob; k < oe; k += NB) {
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
            const int n_in = oe - (k + nb);            // columns of this outer block still to the right
            const int nb2 = n_in > 0 ? (n_in < NB ? n_in : NB) : 0;
            // Critical path first: the next diagonal block only needs the top nb2 rows of the panel.  Solve those,
            // update that one block, and hand it to the look-ahead factorisation; everything else of this panel
            // (the rest of the TRSM, the split, the other updates) then runs underneath the next potrf.
            GemmArgs t{};  // TRSM as GEMM: P <- P * (L_kk^-1)^T, in place; row ranges are independent
            t.B = X + (long long)k * ld + k;
            t.N = nb; t.Kd = nb; t.lda = t.ldb = t.ldc = ld;
            t.alpha = 1.f; t.beta = 0.f;
            int rc;
            if (nb2 > 0) {
                // top rows of the TRSM + update of the next diagonal block: one 10-CTA launch (panel_top_kernel)
                panel_top_kernel<<<10, 256, PT_SMEM, st>>>(A, X, ld, k, nb, nb2);
                rc = check_launch("panel_top");
                if (rc) return rc;
                if (la.ok) {
                    if (cudaEventRecord(la.panel_ready, st) != cudaSuccess) return QT_ERR_CUDA;
                    if (cudaStreamWaitEvent(la.side, la.panel_ready, 0) != cudaSuccess) return QT_ERR_CUDA;
                    potrf_inv_kernel<<<1, 256, smem, la.side>>>(A, X, ld, nb2, k + nb, info);
                    rc = check_launch("potrf_inv(look-ahead)");
                    if (rc) return rc;
                    if (cudaEventRecord(la.potrf_done, la.side) != cudaSuccess) return QT_ERR_CUDA;
                    potrf_ahead = true;
                }
                // the solved top rows go from X to A, together with their tf32 split
                panel_top_commit_kernel<<<8, 256, 0, st>>>(A, X, ld, k, nb, nb2, Lh, Ll);
                rc = check_launch("panel_top_commit");
                if (rc) return rc;
            }
            const long long poff = (long long)(k + nb) * ld + k;
            if (rem > nb2) {     // rest of the TRSM; its epilogue also writes the tf32 split of the solved rows
                float* Pr = P + (long long)nb2 * ld;
                t.A = Pr; t.C = Pr; t.M = rem - nb2;
                t.split_hi = Lh + poff + (long long)nb2 * ld;
                t.split_lo = Ll + poff + (long long)nb2 * ld;
                t.ld_split = ld;
                rc = sgemm(true, t, 1, st);
                if (rc) return rc;
            }
            if (n_in <= 0) continue;
            // inner update (FFMA), only the columns of this outer block, ONE launch:
            //   A[rows below the next diagonal block, (k+nb) .. oe) -= P_rest * P[(k+nb) .. oe, :]^T
            // (column block 0 is the old "s1" against the top rows, the rest the old lower-tile "s2")
            if (rem > nb2) {
                float* Pr = P + (long long)nb2 * ld;
                GemmArgs s12{};
                s12.A = Pr; s12.B = P; s12.C = A + (long long)(k + nb + nb2) * ld + (k + nb);
                s12.M = rem - nb2; s12.N = n_in; s12.Kd = nb; s12.lda = s12.ldb = s12.ldc = ld;
                s12.alpha = -1.f; s12.beta = 1.f; s12.lower_tiles_only = 1; s12.tri_row_offset = nb2;
                rc = sgemm(true, s12, 1, st);
                if (rc) return rc;
            }
        }
        // deferred update of everything right of this outer block: A[oe:, oe:] -= L[oe:, ob:oe] L[oe:, ob:oe]^T,
        // in three parts: (a) the next panel column - its diagonal block goes to the look-ahead factorisation;
        // (b) the other columns of the NEXT outer block; (c) everything beyond it, on the `far` stream, underneath
        // the next outer block's latency-bound panel loop.  Each element still receives its updates in outer-block
        // order (the far stream is ordered, and the next (a)/(b) wait for it), so the result is deterministic.
        const int remo = K - oe;
        if (remo <= 0) break;
        if (far_pending) {
            if (cudaStreamWaitEvent(st, la.far_done, 0) != cudaSuccess) return QT_ERR_CUDA;
            far_pending = false;
        }
        diag_fix_pre_kernel<<<(remo + 7) / 8, 256, 0, st>>>(A, ld, oe, remo, ob, oe - ob, dfix);
        int rc = check_launch("diag_fix_pre");
        if (rc) return rc;
        tgemm::Problem p;
        p.A = {Lh, Ll, K, oe, ld};
        p.B = p.A;
        p.C = A; p.c_rows = K; p.c_cols = K; p.ldc = ld;
        p.Kd = oe - ob;
        p.a_col0 = p.b_col0 = ob;
        p.negate = true; p.accumulate = true; p.lower_tiles_only = true;
        const int nb2 = remo < NB ? remo : NB;
        p.M = remo; p.N = nb2;                                           // (a)
        p.a_row0 = p.b_row0 = p.c_row0 = p.c_col0 = oe;
        rc = tgemm::launch(p, st);
        if (rc) return rc;
        diag_fix_post_kernel<<<1, 256, 0, st>>>(A, ld, oe, nb2, dfix);
        rc = check_launch("diag_fix_post");
        if (rc) return rc;
        if (la.ok && remo > nb2) {
            if (cudaEventRecord(la.panel_ready, st) != cudaSuccess) return QT_ERR_CUDA;
            if (cudaStreamWaitEvent(la.side, la.panel_ready, 0) != cudaSuccess) return QT_ERR_CUDA;
            potrf_inv_kernel<<<1, 256, smem, la.side>>>(A, X, ld, nb2, oe, info);
            rc = check_launch("potrf_inv(look-ahead)");
            if (rc) return rc;
            if (cudaEventRecord(la.potrf_done, la.side) != cudaSuccess) return QT_ERR_CUDA;
            potrf_ahead = true;
        }
        const int nextw = remo < OB ? remo : OB;                         // width of the next outer block
        if (nextw > nb2) {                                               // (b)
            p.M = remo - nb2; p.N = nextw - nb2;
            p.a_row0 = p.b_row0 = p.c_row0 = p.c_col0 = oe + nb2;
            rc = tgemm::launch(p, st);
            if (rc) return rc;
            diag_fix_post_kernel<<<(nextw - nb2 + 255) / 256, 256, 0, st>>>(A, ld, oe + nb2, nextw - nb2, dfix);
            rc = check_launch("diag_fix_post");
            if (rc) return rc;
        }
        if (remo > nextw) {                                              // (c)
            cudaStream_t fs = la.ok ? la.far : st;
            if (la.ok) {
                if (cudaEventRecord(la.far_ready, st) != cudaSuccess) return QT_ERR_CUDA;
                if (cudaStreamWaitEvent(fs, la.far_ready, 0) != cudaSuccess) return QT_ERR_CUDA;
            }
            p.M = remo - nextw; p.N = remo - nextw;
            p.a_row0 = p.b_row0 = p.c_row0 = p.c_col0 = oe + nextw;
            // the far update runs underneath the next outer block's panel loop: leave SMs for its 1-10 CTA kernels
            // (measured: potrf_inv waited ~220 us per outer block for a 148-CTA far update to drain)
            p.reserve_sms = la.ok ? 16 : 0;
            rc = tgemm::launch(p, fs);
            p.reserve_sms = 0;
            if (rc) return rc;
            diag_fix_post_kernel<<<(remo - nextw + 255) / 256, 256, 0, fs>>>(A, ld, oe + nextw, remo - nextw, dfix);
            rc = check_launch("diag_fix_post");
            if (rc) return rc;
            if (la.ok) {
                if (cudaEventRecord(la.far_done, fs) != cudaSuccess) return QT_ERR_CUDA;
                far_pending = true;
            }
        }
    }
    if (far_pending && cudaStreamWaitEvent(st, la.far_done, 0) != cudaSuccess) return QT_ERR_CUDA;
    return QT_OK;
}

// Triangular inverse with the block merges (levels >= 256) on the tensor cores.  Per pair [[A,0],[C,B]]:
//   P = split(X11^T) (upper), S = P C^T, X21 = -X22 S^T - every product contracts over operand columns.
// Dh/Dl hold the operand splits: diagonal positions = P / X22, the (a0, b0) off-diagonal position = S.
static int trtri_lower_tc(const float* L, float* X, float* W, const float* Lh, const float* Ll, float* Dh, float* Dl,
                          int K, int ld, cudaStream_t st) {
    for (long long s = NB; s < K; s *= 2) {
        const long long nblk = (K + s - 1) / s;
        const long long npairs = nblk / 2;
        if (npairs == 0) break;
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
            int rc;
            if (s < 256) {
                GemmArgs g1{};  // T = C * A^-1
                g1.A = L + b0 * ld + a0; g1.B = X + a0 * ld + a0; g1.C = W + b0 * ld + a0;
                g1.M = (int)sB; g1.N = (int)s; g1.Kd = (int)s; g1.lda = g1.ldb = g1.ldc = ld;
                g1.alpha = 1.f; g1.beta = 0.f; g1.b_lower_tri = 1;
                g1.strideA = g1.strideB = g1.strideC = stride;
                rc = sgemm(false, g1, (int)batch, st);
                if (rc) return rc;
                GemmArgs g2{};  // X[C] = -B^-1 * T
                g2.A = X + b0 * ld + b0; g2.B = W + b0 * ld + a0; g2.C = X + b0 * ld + a0;
                g2.M = (int)sB; g2.N = (int)s; g2.Kd = (int)sB; g2.lda = g2.ldb = g2.ldc = ld;
                g2.alpha = -1.f; g2.beta = 0.f; g2.a_lower_tri = 1;
                g2.strideA = g2.strideB = g2.strideC = stride;
                rc = sgemm(false, g2, (int)batch, st);
                if (rc) return rc;
                continue;
            }
            const int is = (int)s, isB = (int)sB, ia0 = (int)a0, ib0 = (int)b0, step = (int)(2 * s);
            // P = split(X11^T) at (a0, a0); X22 split at (b0, b0)
            rc = split_block(true, X + a0 * ld + a0, ld, Dh + a0 * ld + a0, Dl + a0 * ld + a0, ld, is, is, 1, (int)batch,
                             stride, stride, st);
            if (rc) return rc;
            rc = split_block(false, X + b0 * ld + b0, ld, Dh + b0 * ld + b0, Dl + b0 * ld + b0, ld, isB, isB, 1,
                             (int)batch, stride, stride, st);
            if (rc) return rc;
            tgemm::Problem g;   // S[i][l] = sum_k P[i][k] C[l][k]   -> W at (a0, b0)
            g.A = {Dh, Dl, K, K, ld};
            g.B = {Lh, Ll, K, K, ld};
            g.C = W; g.c_rows = K; g.c_cols = K; g.ldc = ld;
            g.M = is; g.N = isB; g.Kd = is;
            g.a_row0 = ia0; g.a_col0 = ia0; g.b_row0 = ib0; g.b_col0 = ia0; g.c_row0 = ia0; g.c_col0 = ib0;
            g.batch = (int)batch;
            g.a_sr = g.a_sc = g.b_sr = g.b_sc = g.c_sr = g.c_sc = step;
            g.a_tri = 2;
            rc = tgemm::launch(g, st);
            if (rc) return rc;
            rc = split_block(false, W + a0 * ld + b0, ld, Dh + a0 * ld + b0, Dl + a0 * ld + b0, ld, is, isB, 0, (int)batch,
                             stride, stride, st);
            if (rc) return rc;
            tgemm::Problem h;   // X21[j][i] = -sum_l X22[j][l] S[i][l]   -> X at (b0, a0)
            h.A = {Dh, Dl, K, K, ld};
            h.B = {Dh, Dl, K, K, ld};
            h.C = X; h.c_rows = K; h.c_cols = K; h.ldc = ld;
            h.M = isB; h.N = is; h.Kd = isB;
            h.a_row0 = ib0; h.a_col0 = ib0; h.b_row0 = ia0; h.b_col0 = ib0; h.c_row0 = ib0; h.c_col0 = ia0;
            h.batch = (int)batch;
            h.a_sr = h.a_sc = h.b_sr = h.b_sc = h.c_sr = h.c_sc = step;
            h.a_tri = 1;
            h.negate = true;
            rc = tgemm::launch(h, st);
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
    const size_t row_bytes = (size_t)K * sizeof(float);
    if (!(K & 3) && row_bytes <= 200 * 1024) {
        // once per chain: set on every call (function attributes are per device)
        if (cudaFuncSetAttribute(gather_flip_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes) !=
            cudaSuccess) { set_last_error("gather_flip smem attr", cudaErrorInvalidValue); return QT_ERR_CUDA; }
        gather_flip_row_kernel<<<K, 512, row_bytes, st>>>(H, perm, dead, damp_scratch, Hf, K);
        return check_launch("gather_flip_row");
    }
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

// Tensor-core variant of qt_gptq_hinv_factor: same contract plus four more K x K fp32 workspaces (tf32 splits of
// L and of the inverse blocks).  K % 256 == 0.  stages: bit 0 = Cholesky trailing updates on the tensor cores,
// bit 1 = triangular-inverse merges on the tensor cores (3 = both; the other stage runs the FFMA GEMMs).
int qt_gptq_hinv_factor_tc(float* A, float* X, float* W, float* Lh, float* Ll, float* Dh, float* Dl, int K, int* info,
                           int stages, void* stream) {
    if (!A || !X || !W || !Lh || !Ll || !Dh || !Dl || !info || K <= 0 || (K & 255)) return QT_ERR_INVALID;
    if (((uintptr_t)A | (uintptr_t)X | (uintptr_t)W | (uintptr_t)Lh | (uintptr_t)Ll | (uintptr_t)Dh | (uintptr_t)Dl) & 15)
        return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
    if (e != cudaSuccess) { set_last_error("memset info", e); return QT_ERR_CUDA; }
    int rc;
    if (stages & 1) {
        rc = cholesky_lower_tc(A, X, Lh, Ll, W, K, K, info, st);   // W's first row holds the diagonal fix-ups
    } else {
        rc = cholesky_lower(A, X, K, K, info, st);
        if (!rc && (stages & 2)) rc = split_block(false, A, K, Lh, Ll, K, K, K, 1, 1, 0, 0, st);
    }
    if (rc) return rc;
    rc = (stages & 2) ? trtri_lower_tc(A, X, W, Lh, Ll, Dh, Dl, K, K, st) : trtri_lower(A, X, W, K, K, st);
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
