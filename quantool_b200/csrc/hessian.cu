// GPTQ Hessian accumulation  H += X^T X  on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces UPSTREAM llmcompressor gptq_quantize.py `accumulate_hessian`,
// the forward hook GPTQModifier installs on every Linear; reached from
// ref/src/quantool/methods/llm_compressor/gptq/gptq.py:86 via llmcompressor.oneshot at
// ref/src/quantool/methods/llm_compressor/base.py:159-161.  Upstream keeps a running mean
// (H *= n/(n+b); H += (2/n) X^T X in fp32); here raw fp32 sums are accumulated over all
// batches and qt_hessian_finalize applies the single factor 2/n_samples and mirrors the
// upper triangle, which is the same matrix up to fp32 rounding.
//
// Shape of the work: X is [T, K] bf16 row-major (K contiguous).  Both MMA operands are tiles
// of the SAME matrix, read "MN-major": a TMA box of 64 columns x 64 tokens lands in shared
// memory as 64 token-rows of 128 bytes (SWIZZLE_128B), which is exactly the canonical
// MN-major SW128 UMMA layout, so no transpose of X is ever materialised.
//   D[128 x 256] (fp32, TMEM) += A[128 x 16]^T-view * B[256 x 16]^T-view per tcgen05.mma
// Only tiles that touch the upper triangle are computed (SYRK); tokens are split S ways so
// that tiles x S fills the 148 SMs evenly; partial tiles are added to H with vectorised
// fp32 reductions (red.global.add.v4.f32) that resolve in L2.
//
// Warp roles (192 threads, 1 CTA/SM, persistent over work units):
//   warp 0    TMA producer      (one elected lane)
//   warp 1    TMEM alloc + tcgen05.mma issuer (one elected lane)
//   warps 2-5 epilogue: tcgen05.ld 32x32b -> registers -> red.global.add
// Pipelines: 4-stage smem ring (full/empty mbarriers), 2 TMEM accumulators (512 columns).
// Reproducibility: the token splits of one tile add into H with `red.global.add.f32` in arrival order, so the
// off-diagonal sums are NOT bit-reproducible from run to run (differences at the last fp32 bit; the diagonal, which
// decides the act_order permutation, is accumulated separately in a fixed order and IS deterministic - see
// diag_partial_kernel / diag_total_kernel).  A second-stage ordered reduce would cost S x 0.4 GB of extra traffic
// per launch at K = 14336; the bench's world = 1 parity block (same path twice) reports the resulting code agreement.
#include "tc_ptx.cuh"

namespace qt {
namespace hess {
using namespace qt::tc;

constexpr int BM = 128;          // rows of the H tile  (columns m0.. of X)
constexpr int BN = 256;          // cols of the H tile  (columns n0.. of X)
constexpr int BKT = 64;          // tokens per pipeline stage
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int BOX_BYTES = 64 * 2 * BKT;               // 64 cols x BKT tokens of bf16 = 8 KB
constexpr int A_BYTES = (BM / 64) * BOX_BYTES;        // 16 KB
constexpr int B_BYTES = (BN / 64) * BOX_BYTES;        // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 48 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 192;

// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4, [16,30) LBO>>4 = stride between 64-element MN atoms (one TMA box),
//   [32,46) SBO>>4 = stride between 8-token groups (8 x 128 B), [46,48) version = 1,
//   [61,64) layout = 2 (SWIZZLE_128B)
QT_D uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16 (format 1) or fp16 (format 0), both MN-major, N=256, M=128
constexpr uint32_t make_idesc(bool bf16) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | (1u << 15) | (1u << 16) |
           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct Sched {
    int K, NI, NJ, ntiles, S, nkb;  // nkb = ceil(T / BKT)
    int nunits;
    uint32_t idesc;                 // operand format (bf16 / fp16) is a run-time property of the activations
};

// tile index -> (im, jn).  The upper-triangle tile set {im <= 2 jn + 1} is walked in compact
// super-tiles of ST_I x ST_J tiles (a 2048 x 2048 patch of H): the ~148 tiles in flight at any
// time then touch ~4K columns of X instead of all K, so each wave streams a third of X from HBM
// (K = 14336: 165 GB -> ~50 GB per Hessian) and the rest of the operand re-reads hit L2.
constexpr int ST_I = 16, ST_J = 8;
QT_D void tile_coords(const Sched& s, int t, int& im, int& jn) {
    for (int sj = 0; sj * ST_J < s.NJ; sj++) {
        const int j0 = sj * ST_J, j1 = (j0 + ST_J < s.NJ) ? j0 + ST_J : s.NJ;
        for (int si = 0; si * ST_I < s.NI; si++) {
            const int lo = si * ST_I;
            if (lo > 2 * (j1 - 1) + 1) break;          // whole super-tile below the diagonal band
            for (int j = j0; j < j1; j++) {
                int hi = 2 * j + 1;
                hi = hi < s.NI - 1 ? hi : s.NI - 1;
                hi = hi < lo + ST_I - 1 ? hi : lo + ST_I - 1;
                const int cnt = hi - lo + 1;
                if (cnt <= 0) continue;
                if (t < cnt) { im = lo + t; jn = j; return; }
                t -= cnt;
            }
        }
    }
    im = 0; jn = 0;   // unreachable for t < ntiles
}

__global__ void __launch_bounds__(NTHREADS, 1)
hessian_syrk_kernel(const __grid_constant__ CUtensorMap tmap, float* __restrict__ H, Sched s) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                  // [STAGES]
    uint64_t* empty = bars + STAGES;        // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;    // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
        for (int i = 0; i < STAGES; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int u = blockIdx.x; u < s.nunits; u += gridDim.x) {
                const int split = u / s.ntiles, tile = u - split * s.ntiles;
                int im, jn;
                tile_coords(s, tile, im, jn);
                const int kb0 = (int)((long long)s.nkb * split / s.S), kb1 = (int)((long long)s.nkb * (split + 1) / s.S);
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    uint8_t* sa = smem + stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
#pragma unroll
                    for (int b = 0; b < BM / 64; b++) tma_load_2d(sa + b * BOX_BYTES, &tmap, &full[stage], im * BM + b * 64, kb * BKT);
#pragma unroll
                    for (int b = 0; b < BN / 64; b++) tma_load_2d(sb + b * BOX_BYTES, &tmap, &full[stage], jn * BN + b * 64, kb * BKT);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = s.idesc;
            uint32_t it = 0, li = 0;
            for (int u = blockIdx.x; u < s.nunits; u += gridDim.x, li++) {
                const int split = u / s.ntiles;
                const int kb0 = (int)((long long)s.nkb * split / s.S), kb1 = (int)((long long)s.nkb * (split + 1) / s.S);
                const uint32_t acc = li & 1, aph = (li >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full[stage], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < BKT / UMMA_K; k++) {
                        const uint64_t ad = make_desc(sa + k * UMMA_K * 128, BOX_BYTES, 1024);
                        const uint64_t bd = make_desc(sb + k * UMMA_K * 128, BOX_BYTES, 1024);
                        tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty[stage]);
                }
                tc_commit(&tfull[acc]);
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> red.add into H =====
        const int lg = warp & 3;  // TMEM lane group this warp may access
        uint32_t li = 0;
        for (int u = blockIdx.x; u < s.nunits; u += gridDim.x, li++) {
            const int split = u / s.ntiles, tile = u - split * s.ntiles;
            int im, jn;
            tile_coords(s, tile, im, jn);
            const int kb0 = (int)((long long)s.nkb * split / s.S), kb1 = (int)((long long)s.nkb * (split + 1) / s.S);
            const uint32_t acc = li & 1, aph = (li >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const int row = im * BM + lg * 32 + lane;
            if (kb1 > kb0) {
#pragma unroll 1
                for (int cc = 0; cc < BN / 32; cc++) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + cc * 32, r);
                    const int col0 = jn * BN + cc * 32;
                    if (row < s.K) {
                        float* hp = H + (long long)row * s.K + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (col0 + j < s.K && col0 + j + 3 >= row)   // keep the 4-group if it touches c >= r
                                red_add_v4(hp + j, __uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                           __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// H <- factor * H on the upper triangle, mirrored to the lower triangle (32x32 smem transpose)
__global__ void __launch_bounds__(256) finalize_kernel(float* __restrict__ H, int K, float factor) {
    __shared__ float t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = bi * 32 + r, j = bj * 32 + tx;
        float v = 0.f;
        if (i < K && j < K) {
            if (bi == bj && j < i) {
                v = 0.f;  // filled from the mirror below
            } else {
                v = H[(long long)i * K + j] * factor;
                H[(long long)i * K + j] = v;
            }
        }
        t[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        // write transposed tile: element (j, i) <- t[i][j]
        const int jj = bj * 32 + r, ii = bi * 32 + tx;
        if (jj < K && ii < K) {
            if (bi == bj) {
                if (ii < jj) H[(long long)jj * K + ii] = t[tx][r];
            } else {
                H[(long long)jj * K + ii] = t[tx][r];
            }
        }
    }
}

// ---- exact diagonal -------------------------------------------------------------------------------
// The tensor cores add into the fp32 TMEM accumulator with truncation; over the ~1000-instruction chain of one
// work unit that biases a sum of squares by about -5e-6 relative (measured), enough to reorder near-equal
// diagonal entries - and the act_order permutation is argsort(diag H).  The diagonal is therefore accumulated
// separately with FFMA in round-to-nearest: a deterministic two-stage column reduction (row slabs -> partial
// sums -> fixed-order total), one extra streaming pass over X (~3 % of the SYRK time).
constexpr int DIAG_SLABS = 32;
template <int DT>
__global__ void __launch_bounds__(256) diag_partial_kernel(const void* __restrict__ X, long long T, int K,
                                                           float* __restrict__ partial) {
    __shared__ float sa[8][256 + 8];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 256 + tx * 8;
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 0.f;
    if (c0 < K) {
        const long long r0 = T * blockIdx.y / gridDim.y, r1 = T * (blockIdx.y + 1) / gridDim.y;
        for (long long r = r0 + ty; r < r1; r += 8) {
            float v[8];
            load8<DT>(X, (r * K + c0) / 8, v);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fmaf(v[i], v[i], a[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) sa[ty][tx * 8 + i] = a[i];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < K) {
        float r = sa[0][threadIdx.x];
#pragma unroll
        for (int j = 1; j < 8; j++) r += sa[j][threadIdx.x];
        partial[(long long)blockIdx.y * K + c] = r;
    }
}
__global__ void __launch_bounds__(256) diag_total_kernel(const float* __restrict__ partial, int K, int slabs,
                                                         float* __restrict__ diag) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= K) return;
    float r = 0.f;
    for (int y = 0; y < slabs; y++) r += partial[(long long)y * K + c];
    diag[c] += r;
}
__global__ void __launch_bounds__(256) set_diag_kernel(float* __restrict__ H, int K, const float* __restrict__ diag) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < K) H[(long long)c * K + c] = diag[c];
}

static int g_force_splits = 0;
static int g_reserve_sms = 0;     // SMs left to other streams (NCCL kernels of an overlapping all-reduce)


// ---- packed upper block-triangle (what crosses NVLink) ------------------------------------------------------
// Before `finalize` H holds raw sums only in the 128 x 256 tiles that touch the upper triangle, and the inverse
// factor U is upper-triangular: shipping the K x K square through the all-reduce / broadcast moves twice the bytes
// (822 MB instead of 420 MB at K = 14336).  Row block i (128 rows) is stored from column c0(i) = floor(128 i / align)
// * align to K, blocks back to back; align = 256 for H (the SYRK's tile width), 128 for U.
__device__ __forceinline__ long long tri_block_offset(int i, int K, int align) {
    long long off = 0;
    for (int b = 0; b < i; b++) off += 128LL * (K - ((b * 128) / align) * align);
    return off;
}

template <bool UNPACK>
__global__ void __launch_bounds__(256) tri_pack_kernel(float* __restrict__ M, float* __restrict__ P, int K, int align,
                                                       int zero_below) {
    const int i = blockIdx.x;
    const int c0 = ((i * 128) / align) * align;
    const int w = K - c0;
    const long long off = tri_block_offset(i, K, align);
    const int rows = min(128, K - i * 128);
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        float* mrow = M + (long long)(i * 128 + r) * K;
        float* prow = P + off + (long long)r * w;
        if (((K | c0) & 3) == 0) {
            float4* m4 = reinterpret_cast<float4*>(mrow + c0);
            float4* p4 = reinterpret_cast<float4*>(prow);
            for (int c = threadIdx.x; c < (w >> 2); c += 256) {
                if (UNPACK) m4[c] = p4[c]; else p4[c] = m4[c];
            }
            if (UNPACK && zero_below) {
                float4* z4 = reinterpret_cast<float4*>(mrow);
                for (int c = threadIdx.x; c < (c0 >> 2); c += 256) z4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int c = threadIdx.x; c < w; c += 256) {
                if (UNPACK) mrow[c0 + c] = prow[c]; else prow[c] = mrow[c0 + c];
            }
            if (UNPACK && zero_below)
                for (int c = threadIdx.x; c < c0; c += 256) mrow[c] = 0.f;
        }
    }
}

}  // namespace hess
}  // namespace qt

using namespace qt;
using namespace qt::hess;

extern "C" {

// debug/tuning knob: force the token split count (0 = heuristic)
int qt_hessian_set_splits(int s) { g_force_splits = s; return QT_OK; }
int qt_hessian_reserve_sms(int n) { g_reserve_sms = n > 0 ? n : 0; return QT_OK; }

// H[K,K] (fp32, zero-initialised by the caller before the first batch) += X^T X over the upper
// triangle tiles.  X: [T, K] bf16 or fp16 row-major (the model's activation dtype, fed to the tensor cores as it
// is: upstream casts the same values to fp32, so no precision is dropped on the way in).  K % 8 == 0.
int qt_hessian_accumulate(const void* X, int dtype, int64_t T, int K, float* H, void* stream) {
    if (dtype == QT_F32) {
        set_last_error("qt_hessian_accumulate: fp32 activations are not supported (bf16 / fp16 tensor-core operands)",
                       cudaSuccess);
        return QT_ERR_UNSUPPORTED;
    }
    if (dtype != QT_BF16 && dtype != QT_F16) return QT_ERR_INVALID;
    if (T < 0 || K <= 0 || (K & 7) || !H) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;     // an empty batch adds nothing (X may be a null pointer then)
    if (!X) return QT_ERR_INVALID;
    if (((uintptr_t)X & 15) || ((uintptr_t)H & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_last_error("cuTensorMapEncodeTiled entry point", cudaErrorUnknown); return QT_ERR_CUDA; }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)T};
    const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)BKT};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tmap, dtype == QT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(X), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue); return QT_ERR_CUDA; }

    Sched s;
    s.idesc = make_idesc(dtype == QT_BF16);
    s.K = K;
    s.NI = (K + BM - 1) / BM;
    s.NJ = (K + BN - 1) / BN;
    s.ntiles = 0;
    for (int j = 0; j < s.NJ; j++) { int c = 2 * j + 2; s.ntiles += c < s.NI ? c : s.NI; }
    s.nkb = (int)((T + BKT - 1) / BKT);
    int dev = 0, nsm = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // token splits.  Measured on B200 (K = 14336, T = 262144): 53.5 ms at S = 1, 49.2 at S = 8, 47.5 at
    // S = 16 - the CTAs of one split share a token slab, and the slab has to be small enough
    // (<= ~512 MB of X) for its re-reads across waves to be served by L2.  Above that floor pick the
    // split count with the best wave quantisation (+6 k-blocks of per-unit overhead).
    int S = 1;
    if (g_force_splits > 0) {
        S = g_force_splits;
    } else {
        const double slab_bytes = (double)T * K * 2.0;
        int smin = (int)((slab_bytes + 536870911.0) / 536870912.0);
        if (smin < 1) smin = 1;
        int smax = s.nkb / 16 > 1 ? s.nkb / 16 : 1;
        if (smin > smax) smin = smax;
        if (smax > 2 * smin + 8) smax = 2 * smin + 8;
        double best = 1e30;
        for (int c = smin; c <= smax; c++) {
            const long long units = (long long)s.ntiles * c;
            const long long waves = (units + nsm - 1) / nsm;
            const double eff_time = (double)waves * ((double)s.nkb / c + 6.0);
            if (eff_time < best * 0.995) { best = eff_time; S = c; }
        }
    }
    if (S > s.nkb) S = s.nkb;
    s.S = S;
    s.nunits = s.ntiles * S;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(hessian_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("hessian smem attr", e); return QT_ERR_CUDA; }
        attr_set = true;
    }
    // The persistent grid normally takes every SM; with an all-reduce of a previous Hessian in flight its NCCL CTAs
    // could not be placed next to the SYRK's (register file), so the caller may reserve a few SMs for them.
    int avail = nsm - g_reserve_sms;
    if (avail < nsm / 2) avail = nsm / 2;
    const int grid = s.nunits < avail ? s.nunits : avail;
    hessian_syrk_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(tmap, H, s);
    return check_launch("hessian_syrk");
}

int qt_hessian_finalize(float* H, int K, float factor, void* stream) {
    if (!H || K <= 0) return QT_ERR_INVALID;
    const int nb = (K + 31) / 32;
    dim3 grid(nb, nb);
    finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(H, K, factor);
    return check_launch("hessian_finalize");
}

// diag[c] += sum_t X[t][c]^2 in fp32 round-to-nearest (deterministic); scratch: 32 * K floats
int qt_hessian_diag_accumulate(const void* X, int dtype, int64_t T, int K, float* diag, float* scratch, void* stream) {
    if (dtype != QT_BF16 && dtype != QT_F16) return QT_ERR_INVALID;
    if (T < 0 || K <= 0 || (K & 7) || !diag || !scratch) return QT_ERR_INVALID;
    if (T == 0) return QT_OK;
    if (!X || ((uintptr_t)X & 15)) return QT_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = T < DIAG_SLABS ? (int)T : DIAG_SLABS;
    dim3 grid((K + 255) / 256, slabs);
    if (dtype == QT_BF16) diag_partial_kernel<QT_BF16><<<grid, 256, 0, st>>>(X, T, K, scratch);
    else diag_partial_kernel<QT_F16><<<grid, 256, 0, st>>>(X, T, K, scratch);
    int rc = check_launch("hessian_diag_partial");
    if (rc) return rc;
    diag_total_kernel<<<(K + 255) / 256, 256, 0, st>>>(scratch, K, slabs, diag);
    return check_launch("hessian_diag_total");
}

// H[c][c] = diag[c] (raw sums; call before qt_hessian_finalize and before any cross-rank reduction of H)
int qt_hessian_set_diagonal(float* H, int K, const float* diag, void* stream) {
    if (!H || !diag || K <= 0) return QT_ERR_INVALID;
    set_diag_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(H, K, diag);
    return check_launch("hessian_set_diagonal");
}

// Packed upper block-triangle of a K x K fp32 matrix (see tri_pack_kernel).  align: 256 for raw Hessian sums,
// 128 for the upper-triangular factor U.  qt_tri_packed_elems gives the element count of the packed buffer.
int64_t qt_tri_packed_elems(int K, int align) {
    if (K <= 0 || (align != 128 && align != 256)) return -1;
    long long n = 0;
    for (int b = 0; b * 128 < K; b++) {
        const int rows = (K - b * 128) < 128 ? (K - b * 128) : 128;
        n += (long long)rows * (K - ((b * 128) / align) * align);
    }
    return n;
}

int qt_tri_pack(const float* M, int K, int align, float* packed, void* stream) {
    if (!M || !packed || K <= 0 || (align != 128 && align != 256)) return QT_ERR_INVALID;
    dim3 grid((K + 127) / 128, 16);
    tri_pack_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(const_cast<float*>(M), packed, K, align, 0);
    return check_launch("tri_pack");
}

int qt_tri_unpack(const float* packed, int K, int align, int zero_below, float* M, void* stream) {
    if (!M || !packed || K <= 0 || (align != 128 && align != 256)) return QT_ERR_INVALID;
    dim3 grid((K + 127) / 128, 16);
    tri_pack_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(M, const_cast<float*>(packed), K, align, zero_below);
    return check_launch("tri_unpack");
}

}  // extern "C"
