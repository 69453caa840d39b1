// fp32 SIMT GEMM used by the Cholesky chain and the GPTQ lazy-batch update, where the
// reference keeps everything in fp32 (GPTQ_PRECISION = torch.float32) and
// north_star asks that the trailing updates stay in fp32.
//
//   C[M,N] = alpha * A[M,Kd] * op(B) + beta * C        (row-major, leading dims lda/ldb/ldc)
//   B_NK = false : B is [Kd, N] (n contiguous)          "NN"
//   B_NK = true  : B is [N, Kd] (k contiguous)          "NT"  (C += A * B^T)
//
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile of packed f32x2 FMAs (two 4-wide halves per
// dimension so every shared-memory read is a conflict-free LDS.128), global->register
// prefetch of the next k-slab overlapped with the FMAs of the current one, two shared
// buffers, one __syncthreads per slab.  Triangular structure is skipped at tile granularity.
#pragma once
#include "common.cuh"

namespace qt {

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    int M, N, Kd;
    int lda, ldb, ldc;
    float alpha, beta;
    long long strideA, strideB, strideC;  // batch strides (elements); batch = gridDim.z
    int lower_tiles_only;  // 1: skip tiles entirely above the diagonal (SYRK into a lower triangle)
    int a_lower_tri;       // 1: A[m][k] == 0 for k > m  -> k range ends at the tile's last row
    int b_lower_tri;       // 1: B[k][n] == 0 for k < n  -> k range starts at the tile's first col
    int tri_row_offset;    // rows of this call start at this row of the triangular structure (row slices)
    // Kd = 128 "NT" kernel only: also write the tf32 split of the result (what the tensor-core updates read as
    // operands) - hi/lo at the same (row, column) offsets as C inside buffers with leading dimension ld_split
    float* split_hi;
    float* split_lo;
    int ld_split;
};

constexpr int GBM = 128, GBN = 128, GBK = 16;
constexpr int GLD = GBM + 4;  // smem leading dim (keeps float4 alignment)

template <bool B_NK>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][GBK][GLD];
    __shared__ __align__(16) float Bs[2][GBK][GLD];

    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    if (g.lower_tiles_only && n0 > m0 + g.tri_row_offset + GBM - 1) return;
    const float* __restrict__ A = g.A + (long long)blockIdx.z * g.strideA;
    const float* __restrict__ B = g.B + (long long)blockIdx.z * g.strideB;
    float* __restrict__ C = g.C + (long long)blockIdx.z * g.strideC;

    int k_begin = 0, k_end = g.Kd;
    if (g.a_lower_tri) { const int e = m0 + g.tri_row_offset + GBM; k_end = e < k_end ? e : k_end; }
    if (g.b_lower_tri) { k_begin = (n0 / GBK) * GBK; }

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;

    // global->smem mapping for a k-contiguous operand tile (128 rows x 16 k): 512 float4,
    // thread t handles rows r = t/4 and r+64, k-quad kq = t%4
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    // mapping for an n-contiguous B slab (16 k x 128 n): 512 float4, thread t handles
    // k = t/32 and k+8, n-quad = (t%32)*4
    const int bk = tid >> 5, bn = (tid & 31) * 4;

    float4 ra[2], rb[2];
    auto load_slab = [&](int k0) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int m = m0 + lr + 64 * h, k = k0 + lk;
            ra[h] = (m < g.M && k < k_end) ? *reinterpret_cast<const float4*>(A + (long long)m * g.lda + k)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (B_NK) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int n = n0 + lr + 64 * h, k = k0 + lk;
                rb[h] = (n < g.N && k < k_end) ? *reinterpret_cast<const float4*>(B + (long long)n * g.ldb + k)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int k = k0 + bk + 8 * h, n = n0 + bn;
                rb[h] = (k < k_end && n < g.N) ? *reinterpret_cast<const float4*>(B + (long long)k * g.ldb + n)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto store_slab = [&](int buf) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lr + 64 * h;
            As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
            As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
        }
        if (B_NK) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = lr + 64 * h;
                Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
                Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
            }
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++)
                *reinterpret_cast<float4*>(&Bs[buf][bk + 8 * h][bn]) = rb[h];
        }
    };

    // accumulators as packed fp32 pairs (acc2[i][jp] = {acc[i][2jp], acc[i][2jp+1]}): Blackwell's
    // fma.rn.f32x2 (SASS FFMA2) retires two FMAs per issue slot, and this kernel is issue-bound on
    // plain FFMA: 44.6 -> 49-52 TFLOP/s at 8192^3.  (Keeping A pre-duplicated in shared memory to
    // drop the {a,a} register moves was slower - 43 TFLOP/s - the extra LDS cost more than the moves.)
    unsigned long long acc2[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc2[i][j] = 0ull;

    const int nslab = (k_end - k_begin + GBK - 1) / GBK;
    if (nslab > 0) {
        load_slab(k_begin);
        store_slab(0);
        __syncthreads();
        for (int s = 0; s < nslab; s++) {
            const int buf = s & 1;
            if (s + 1 < nslab) load_slab(k_begin + (s + 1) * GBK);
#pragma unroll
            for (int k = 0; k < GBK; k++) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                unsigned long long bp[4];
                asm("mov.b64 %0, {%1, %2};" : "=l"(bp[0]) : "f"(b0.x), "f"(b0.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bp[1]) : "f"(b0.z), "f"(b0.w));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bp[2]) : "f"(b1.x), "f"(b1.y));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bp[3]) : "f"(b1.z), "f"(b1.w));
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    unsigned long long ap;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(ap) : "f"(a[i]));
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(ap), "l"(bp[j]));
                }
            }
            if (s + 1 < nslab) {
                store_slab(buf ^ 1);
                __syncthreads();
            }
        }
    }

    // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4.., 64+tx*4..}
#pragma unroll
    for (int ih = 0; ih < 2; ih++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = m0 + 64 * ih + ty * 4 + i;
            if (m >= g.M) continue;
#pragma unroll
            for (int jh = 0; jh < 2; jh++) {
                const int n = n0 + 64 * jh + tx * 4;
                if (n >= g.N) continue;  // N % 4 == 0 is enforced by the host wrapper
                float* cp = C + (long long)m * g.ldc + n;
                float4 o;
                float v[4];
                asm("mov.b64 {%0, %1}, %2;" : "=f"(v[0]), "=f"(v[1]) : "l"(acc2[4 * ih + i][2 * jh]));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(v[2]), "=f"(v[3]) : "l"(acc2[4 * ih + i][2 * jh + 1]));
                if (g.beta != 0.f) {
                    const float4 c = *reinterpret_cast<const float4*>(cp);
                    o.x = g.alpha * v[0] + g.beta * c.x; o.y = g.alpha * v[1] + g.beta * c.y;
                    o.z = g.alpha * v[2] + g.beta * c.z; o.w = g.alpha * v[3] + g.beta * c.w;
                } else {
                    o.x = g.alpha * v[0]; o.y = g.alpha * v[1]; o.z = g.alpha * v[2]; o.w = g.alpha * v[3];
                }
                *reinterpret_cast<float4*>(cp) = o;
            }
        }
}

// ---- Kd = 128, "NT": the panel products of the blocked Cholesky -------------------------------------------
// The chain's TRSMs and inner SYRK updates are skinny (N = 128..384, inner dimension exactly 128) and sit between
// latency-bound kernels, so what matters is how long ONE tile takes: the slab pipeline above makes eight dependent
// global-memory round trips for Kd = 128 (24-35 us per launch, CUPTI).  Here the whole k range of a 64 x 128 tile
// is staged with one wave of cp.async (both operands k-contiguous, rows padded to 132 floats: LDS.128 along k is
// conflict-free), one barrier, then 4 x 8 outputs per thread straight from shared memory.
constexpr int SK_BM = 64, SK_BN = 128, SK_KD = 128, SK_LD = SK_KD + 4;
constexpr size_t SK_SMEM = (size_t)(SK_BM + SK_BN) * SK_LD * sizeof(float);

template <int UNUSED>   // a template so that the header can be included from several translation units
__global__ void __launch_bounds__(256, 2) sgemm_nt_k128_kernel(GemmArgs g) {
    extern __shared__ __align__(16) float sk_sm[];
    float(*As)[SK_LD] = reinterpret_cast<float(*)[SK_LD]>(sk_sm);
    float(*Bs)[SK_LD] = reinterpret_cast<float(*)[SK_LD]>(sk_sm + SK_BM * SK_LD);
    const int m0 = blockIdx.y * SK_BM, n0 = blockIdx.x * SK_BN;
    if (g.lower_tiles_only && n0 > m0 + g.tri_row_offset + SK_BM - 1) return;
    const int tid = threadIdx.x;
    // (64 + 128) rows x 32 float4: 24 cp.async per thread, all in flight at once; out-of-range rows are zero-filled
    for (int q = tid; q < (SK_BM + SK_BN) * (SK_KD / 4); q += 256) {
        const int r = q >> 5, k4 = (q & 31) * 4;
        const float* src;
        bool ok;
        float* dst;
        if (r < SK_BM) { ok = (m0 + r) < g.M; src = g.A + (long long)(ok ? m0 + r : 0) * g.lda + k4; dst = &As[r][k4]; }
        else { const int n = r - SK_BM; ok = (n0 + n) < g.N; src = g.B + (long long)(ok ? n0 + n : 0) * g.ldb + k4; dst = &Bs[n][k4]; }
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;          // rows 4 ty + i, columns tx + 16 j
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < SK_KD; k += 4) {
        float4 a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const float4*>(&As[4 * ty + i][k]);
#pragma unroll
        for (int j = 0; j < 8; j++) b[j] = *reinterpret_cast<const float4*>(&Bs[tx + 16 * j][k]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int m = m0 + 4 * ty + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int n = n0 + tx + 16 * j;
            if (n >= g.N) continue;
            float* cp = g.C + (long long)m * g.ldc + n;
            const float v = (g.beta != 0.f) ? g.alpha * acc[i][j] + g.beta * (*cp) : g.alpha * acc[i][j];
            *cp = v;
            if (g.split_hi) {
                float h, l;
                tf32_split(v, h, l);
                g.split_hi[(long long)m * g.ld_split + n] = h;
                g.split_lo[(long long)m * g.ld_split + n] = l;
            }
        }
    }
}

// returns QT_OK / QT_ERR_INVALID (alignment contract: dims and leading dims multiples of 4,
// pointers 16-byte aligned).  Empty problems are a no-op.
inline int sgemm(bool b_nk, const GemmArgs& g, int batch, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return QT_OK;
    if ((g.N & 3) || (g.Kd & 3) || (g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return QT_ERR_INVALID;
    if (((uintptr_t)g.A & 15) || ((uintptr_t)g.B & 15) || ((uintptr_t)g.C & 15)) return QT_ERR_INVALID;
    if ((g.strideA & 3) || (g.strideB & 3) || (g.strideC & 3)) return QT_ERR_INVALID;
    if (b_nk && g.Kd == SK_KD && batch == 1 && !g.a_lower_tri && !g.b_lower_tri) {
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute(sgemm_nt_k128_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SK_SMEM) !=
                cudaSuccess)
                return QT_ERR_CUDA;
            attr_set = true;
        }
        dim3 sgrid((g.N + SK_BN - 1) / SK_BN, (g.M + SK_BM - 1) / SK_BM, 1);
        sgemm_nt_k128_kernel<0><<<sgrid, 256, SK_SMEM, st>>>(g);
        return check_launch("sgemm_nt_k128");
    }
    dim3 grid((g.N + GBN - 1) / GBN, (g.M + GBM - 1) / GBM, batch);
    if (b_nk) sgemm_kernel<true><<<grid, 256, 0, st>>>(g);
    else      sgemm_kernel<false><<<grid, 256, 0, st>>>(g);
    return check_launch("sgemm");
}

}  // namespace qt
