// fp32-faithful GEMM on the 5th-gen tensor cores for the inverse-Hessian chain:
//     C (op)= +-A * B^T,   A [M,Kd], B [N,Kd] both row-major = K-major, fp32 in, fp32 out.
//
// Replaces the fp32 cuSOLVER/cuBLAS work behind `torch.linalg.cholesky` / `cholesky_inverse` in UPSTREAM
// llmcompressor gptq_quantize.py, i.e. the trailing SYRK updates of the blocked Cholesky and the
// block merges of the triangular inverse, which were FFMA GEMMs at ~50 TFLOP/s in the first version.
//
// tcgen05 kind::tf32 with operand splitting ("3xTF32"): x ~= hi + lo, hi = tf32(x), lo = tf32(x - hi) (about 22 mantissa bits), and
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi   (dropped lo*lo term ~2^-22 relative), fp32 accumulation in TMEM.
// Callers hand in the split operands (they are produced once per panel / block by the split kernels and reused
// by many tiles).  Pipeline: TMA-staged operand ring -> one-thread tcgen05.mma issue -> TMEM epilogue; arbitrary inner
// dimension, batches of diagonal blocks, triangular operands (k range trimmed per tile), lower-tiles-only
// SYRK, and two epilogues - TMA store (C =) or TMA reduce-add in L2 (C +=), so C is never read by the SM.
//
// Accuracy note (measured): the tensor cores add into the fp32 TMEM accumulator with truncation, so a long
// accumulation chain drifts: with one 3 x Kd/8-instruction chain per tile the inverse factor at K = 14336 was
// 2.8e-5 away from the FFMA result, and the error is linear in the chain length (8.1e-6 / 4.7e-6 / 2.4e-6 /
// 1.3e-6 at 1024 / 512 / 256 / 128).  The chain is therefore cut every `max_chain` = 128 k: each partial sum goes
// to C through the TMA (first one stored, the rest reduce-added in L2 with round-to-nearest, strictly in order, so
// the result is deterministic).  At 128 the factor is within 2x of the FFMA chain's distance to fp64 (K = 4096:
// 5.0e-7 vs 2.8e-7) for 9 % of the chain time.
//
// Tile 128 x 256, k chunks of 32 floats (one 128-byte swizzle row); stage = A_hi,A_lo (2 x 16 KB) + B_hi,B_lo
// (2 x 32 KB) = 96 KB, 2 stages; 2 TMEM accumulators (2 x 256 columns); warp 0 = TMA producer, warp 1 = MMA
// issuer, warps 2-5 = epilogue (32 rows each, 32 x 32 sub-tiles through swizzled staging + TMA).
#include "tgemm.cuh"

#include "tc_ptx.cuh"

namespace qt {
namespace tgemm {
using namespace qt::tc;

constexpr int BM = 128, BN = 256, KC = 32;
constexpr int UMMA_K = 8;  // tf32
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * KC * 4;
constexpr int B_BYTES = BN * KC * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
constexpr int EPI_BYTES = 4 * 2 * 4096;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
constexpr int NTHREADS = 192;

// kind::tf32: D=f32 (bit 4), A=B=TF32 (2 at bits 7,10), negate A (bit 13), A and B both K-major
constexpr uint32_t make_idesc(bool negate) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((negate ? 1u : 0u) << 13) | ((uint32_t)(BN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

struct Args {
    int m_tiles, n_tiles, batch, kchunks;
    int M, N;
    int a_row0, a_col0, b_row0, b_col0, c_row0, c_col0;
    int a_sr, a_sc, b_sr, b_sc, c_sr, c_sc;
    int accumulate, lower_only, a_tri, b_tri;
    int chain_kc;   // k chunks accumulated in TMEM before the partial sum is flushed to C (see Problem::max_chain)
    uint32_t idesc;
    int single;     // hi * hi only
    int bf16;       // operands are bf16 (kind::f16), k chunk = 64 elements
    int kce;        // elements per k chunk (one 128-byte swizzle row): 32 fp32 or 64 bf16
    const float* E; // reduction epilogue (see Problem::E)
    long long lde;
    double* loss;
};

struct Tile {
    int b, tm, tn, kc0, kc1;
};

// decode tile index t; false = nothing to do for this tile (above the diagonal, or empty k range)
QT_D bool decode(const Args& a, int t, Tile& o) {
    const int per = a.m_tiles * a.n_tiles;
    o.b = t / per;
    const int r = t - o.b * per;
    o.tn = r / a.m_tiles;       // n-major: consecutive CTAs share the same B rows (L2 reuse)
    o.tm = r - o.tn * a.m_tiles;
    if (a.lower_only && o.tn * BN > o.tm * BM + BM - 1) return false;
    const int KCe = a.kce;
    int k0 = 0, k1 = a.kchunks * KCe;
    if (a.a_tri == 1) k1 = min(k1, o.tm * BM + BM);
    if (a.a_tri == 2) k0 = max(k0, o.tm * BM);
    if (a.b_tri == 1) k1 = min(k1, o.tn * BN + BN);
    if (a.b_tri == 2) k0 = max(k0, o.tn * BN);
    o.kc0 = k0 / KCe;
    o.kc1 = (k1 + KCe - 1) / KCe;
    return o.kc1 > o.kc0;
}

__global__ void __launch_bounds__(NTHREADS, 1)
tgemm_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
             const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
             const __grid_constant__ CUtensorMap map_c, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi = smem + STAGES * STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi + EPI_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tfull = bars + 2 * STAGES;
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = a.batch * a.m_tiles * a.n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_alo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
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
        if (lane == 0) {
            uint32_t it = 0;
            Tile tl;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                if (!decode(a, t, tl)) continue;
                const int ra = a.a_row0 + tl.b * a.a_sr + tl.tm * BM, ca = a.a_col0 + tl.b * a.a_sc;
                const int rb = a.b_row0 + tl.b * a.b_sr + tl.tn * BN, cb = a.b_col0 + tl.b * a.b_sc;
                for (int kc = tl.kc0; kc < tl.kc1; kc++, it++) {   // sub-ranges need no special handling here
                    const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_expect_tx(&full[stage], a.single ? A_BYTES + B_BYTES : STAGE_BYTES);
                    uint8_t* s = smem + stage * STAGE_BYTES;
                    tma_load_2d(s, &map_ahi, &full[stage], ca + kc * a.kce, ra);
                    if (!a.single) tma_load_2d(s + A_BYTES, &map_alo, &full[stage], ca + kc * a.kce, ra);
                    tma_load_2d(s + 2 * A_BYTES, &map_bhi, &full[stage], cb + kc * a.kce, rb);
                    if (!a.single) tma_load_2d(s + 2 * A_BYTES + B_BYTES, &map_blo, &full[stage], cb + kc * a.kce, rb);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, li = 0;
            Tile tl;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                if (!decode(a, t, tl)) continue;
                for (int k0 = tl.kc0; k0 < tl.kc1; k0 += a.chain_kc) {
                    const int k1 = min(k0 + a.chain_kc, tl.kc1);
                    const uint32_t acc = li & 1, aph = (li >> 1) & 1;
                    li++;
                    mbar_wait(&tempty[acc], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    for (int kc = k0; kc < k1; kc++, it++) {
                        const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(&full[stage], ph);
                        tc_fence_after();
                        const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES);
                        const uint32_t sa_lo = sa_hi + A_BYTES;
                        const uint32_t sb_hi = sa_hi + 2 * A_BYTES;
                        const uint32_t sb_lo = sb_hi + B_BYTES;
#pragma unroll
                        for (int k = 0; k < KC / UMMA_K; k++) {
                            // K-major SW128: rows of 128 B, 8-row groups 1 KB apart; k-step = 32 B inside the row
                            const uint64_t ahi = make_desc_sw128(sa_hi + k * UMMA_K * 4, 16, 1024);
                            const uint64_t alo = make_desc_sw128(sa_lo + k * UMMA_K * 4, 16, 1024);
                            const uint64_t bhi = make_desc_sw128(sb_hi + k * UMMA_K * 4, 16, 1024);
                            const uint64_t blo = make_desc_sw128(sb_lo + k * UMMA_K * 4, 16, 1024);
                            if (a.bf16) {          // 32 bytes per step either way: 8 tf32 or 16 bf16
                                tc_mma_bf16(d_tmem, ahi, bhi, a.idesc, (kc > k0 || k > 0) ? 1u : 0u);
                                continue;
                            }
                            tc_mma_tf32(d_tmem, ahi, bhi, a.idesc, (kc > k0 || k > 0) ? 1u : 0u);
                            if (!a.single) {
                                tc_mma_tf32(d_tmem, ahi, blo, a.idesc, 1u);
                                tc_mma_tf32(d_tmem, alo, bhi, a.idesc, 1u);
                            }
                        }
                        tc_commit(&empty[stage]);
                    }
                    tc_commit(&tfull[acc]);
                }
            }
        }
    } else {
        const int lg = warp & 3;
        uint8_t* my = epi + (warp - 2) * 2 * 4096;
        uint32_t li = 0, nstore = 0;
        double wsum = 0.0;      // reduction epilogue: this lane's share of sum(acc * E)
        Tile tl;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            if (!decode(a, t, tl)) continue;
            const int row_local = tl.tm * BM + lg * 32;
            const int crow = a.c_row0 + tl.b * a.c_sr + row_local;
            const int ccol0 = a.c_col0 + tl.b * a.c_sc + tl.tn * BN;
            for (int k0 = tl.kc0; k0 < tl.kc1; k0 += a.chain_kc) {
                const uint32_t acc = li & 1, aph = (li >> 1) & 1;
                li++;
                mbar_wait(&tfull[acc], aph);
                tc_fence_after();
                if (a.E) {
                    // lane = row: 32 consecutive fp32 of E per sub-tile (one 128-byte line per lane), fp32 products
                    // summed per sub-tile, fp64 across sub-tiles
                    if (row_local < a.M) {
#pragma unroll 1
                        for (int cc = 0; cc < BN / 32; cc++) {
                            if (tl.tn * BN + cc * 32 >= a.N) break;
                            uint32_t r[32];
                            tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + cc * 32, r);
                            if (row_local + lane < a.M) {
                                const float* e = a.E + (long long)(crow + lane) * a.lde + ccol0 + cc * 32;
                                const int ncol = min(32, a.N - (tl.tn * BN + cc * 32));
                                float s4 = 0.f;
                                if (ncol == 32) {
#pragma unroll
                                    for (int c = 0; c < 8; c++) {
                                        const float4 v = *reinterpret_cast<const float4*>(e + 4 * c);
                                        s4 = fmaf(__uint_as_float(r[4 * c]), v.x, s4);
                                        s4 = fmaf(__uint_as_float(r[4 * c + 1]), v.y, s4);
                                        s4 = fmaf(__uint_as_float(r[4 * c + 2]), v.z, s4);
                                        s4 = fmaf(__uint_as_float(r[4 * c + 3]), v.w, s4);
                                    }
                                } else {
#pragma unroll
                                    for (int c = 0; c < 32; c++)
                                        if (c < ncol) s4 = fmaf(__uint_as_float(r[c]), e[c], s4);
                                }
                                wsum += (double)s4;
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(&tempty[acc]);
                    continue;
                }
                // partial sums of one tile are applied in order: the previous flush must have completed
                const bool first = (k0 == tl.kc0);
                if (!first && lane == 0) bulk_wait<0>();
                const bool add = a.accumulate || !first;
                if (row_local < a.M) {
#pragma unroll 1
                    for (int cc = 0; cc < BN / 32; cc++) {
                        if (tl.tn * BN + cc * 32 >= a.N) break;
                        uint32_t r[32];
                        tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + cc * 32, r);
                        uint8_t* buf = my + (nstore & 1) * 4096;
                        nstore++;
                        // the buffer used two stores ago must have been read by the TMA engine
                        if (lane == 0) bulk_wait_read<1>();
                        __syncwarp();
                        // row = lane (128 B), 16-byte chunk c stored at c ^ (lane & 7): SWIZZLE_128B, conflict-free
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const uint4 v = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
                            *reinterpret_cast<uint4*>(buf + lane * 128 + ((c ^ (lane & 7)) << 4)) = v;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (add) tma_reduce_add_2d(&map_c, buf, ccol0 + cc * 32, crow);
                            else tma_store_2d(&map_c, buf, ccol0 + cc * 32, crow);
                            bulk_commit();
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty[acc]);
            }
        }
        if (lane == 0) bulk_wait<0>();
        if (a.E) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_down_sync(0xffffffffu, wsum, o);
            if (lane == 0 && wsum != 0.0) atomicAdd(a.loss, wsum);
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

int launch(const Problem& p, cudaStream_t st) {
    if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return QT_OK;
    const int kce = p.bf16 ? 2 * KC : KC;
    const int esz = p.bf16 ? 2 : 4;
    if (p.bf16 && !p.single) return QT_ERR_INVALID;
    if (p.Kd <= 0 || (p.Kd % kce) || !p.A.hi || !p.B.hi) return QT_ERR_INVALID;
    if (!p.single && (!p.A.lo || !p.B.lo)) return QT_ERR_INVALID;
    if (p.E ? (!p.loss || (p.lde & 3) || ((uintptr_t)p.E & 15) || p.batch != 1) : !p.C) return QT_ERR_INVALID;
    if (p.batch > 1 && ((p.M % BM) || (p.N % BN))) return QT_ERR_INVALID;
    if ((p.A.ld & (p.bf16 ? 7 : 3)) || (p.B.ld & (p.bf16 ? 7 : 3)) || (!p.E && (p.ldc & 3))) return QT_ERR_INVALID;
    const CUtensorMapDataType F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapDataType OPT = p.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : F32;
    const CUtensorMapSwizzle SW = CU_TENSOR_MAP_SWIZZLE_128B;
    // stores / reduce-adds are clipped to the last batch's sub-block
    int crows = p.c_row0 + (p.batch - 1) * p.c_sr + p.M, ccols = p.c_col0 + (p.batch - 1) * p.c_sc + p.N;
    if (!p.E && (crows > p.c_rows || ccols > p.c_cols)) return QT_ERR_INVALID;
    const float* a_lo = p.single ? p.A.hi : p.A.lo;      // unused maps still need a valid address
    const float* b_lo = p.single ? p.B.hi : p.B.lo;
    float* c_ptr = p.E ? const_cast<float*>(p.E) : p.C;
    const uint64_t c_ld = p.E ? (uint64_t)p.lde : (uint64_t)p.ldc;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mc;
    bool ok = make_map_2d(&ma_hi, OPT, p.A.hi, p.A.cols, p.A.rows, (uint64_t)p.A.ld * esz, kce, BM, SW) &&
              make_map_2d(&ma_lo, OPT, a_lo, p.A.cols, p.A.rows, (uint64_t)p.A.ld * esz, kce, BM, SW) &&
              make_map_2d(&mb_hi, OPT, p.B.hi, p.B.cols, p.B.rows, (uint64_t)p.B.ld * esz, kce, BN, SW) &&
              make_map_2d(&mb_lo, OPT, b_lo, p.B.cols, p.B.rows, (uint64_t)p.B.ld * esz, kce, BN, SW) &&
              make_map_2d(&mc, F32, c_ptr, ccols, crows, c_ld * 4, 32, 32, SW);
    if (!ok) { set_last_error("tgemm tensor maps", cudaErrorInvalidValue); return QT_ERR_CUDA; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("tgemm smem attr", e); return QT_ERR_CUDA; }
        attr_set = true;
    }
    Args a;
    a.m_tiles = (p.M + BM - 1) / BM;
    a.n_tiles = (p.N + BN - 1) / BN;
    a.batch = p.batch;
    a.kchunks = p.Kd / kce;
    a.kce = kce; a.bf16 = p.bf16 ? 1 : 0;
    a.M = p.M; a.N = p.N;
    a.a_row0 = p.a_row0; a.a_col0 = p.a_col0; a.b_row0 = p.b_row0; a.b_col0 = p.b_col0;
    a.c_row0 = p.c_row0; a.c_col0 = p.c_col0;
    a.a_sr = p.a_sr; a.a_sc = p.a_sc; a.b_sr = p.b_sr; a.b_sc = p.b_sc; a.c_sr = p.c_sr; a.c_sc = p.c_sc;
    a.accumulate = p.accumulate; a.lower_only = p.lower_tiles_only; a.a_tri = p.a_tri; a.b_tri = p.b_tri;
    a.chain_kc = p.max_chain >= kce ? p.max_chain / kce : 1;
    a.idesc = p.negate ? make_idesc(true) : make_idesc(false);
    // kind::f16 with bf16 operands: format 1 at bits 7 and 10 instead of tf32's 2
    if (p.bf16) a.idesc = (a.idesc & ~((7u << 7) | (7u << 10))) | (1u << 7) | (1u << 10);
    a.single = p.single ? 1 : 0;
    a.E = p.E; a.lde = p.lde; a.loss = p.loss;
    if (p.E) a.chain_kc = a.kchunks;          // one chain per tile: every flush would re-read the E tile
    int dev = 0, nsm = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (long long)a.batch * a.m_tiles * a.n_tiles;
    if (p.reserve_sms > 0 && p.reserve_sms < nsm) nsm -= p.reserve_sms;
    tgemm_kernel<<<(unsigned)(ntiles < nsm ? ntiles : nsm), NTHREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, mc, a);
    return check_launch("tgemm");
}

}  // namespace tgemm
}  // namespace qt

using namespace qt;

extern "C" {

// C (op)= +-A B^T with A = a_hi + a_lo [M,Kd], B = b_hi + b_lo [N,Kd] (tf32 splits, qt_split_tf32), all row-major
// with leading dimensions lda/ldb/ldc.  flags: bit0 negate, bit1 accumulate (C +=), bit2 lower tiles only,
// bits 4-5 a_tri, bits 6-7 b_tri (1 lower, 2 upper).  Kd % 32 == 0; pointers 16-byte aligned.
int qt_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, float* C, int M, int N,
                   int Kd, int lda, int ldb, int ldc, int flags, void* stream) {
    if (!a_hi || !a_lo || !b_hi || !b_lo || !C || M < 0 || N < 0 || Kd <= 0) return QT_ERR_INVALID;
    if (((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo | (uintptr_t)C) & 15) return QT_ERR_INVALID;
    tgemm::Problem p;
    p.A = {a_hi, a_lo, M, Kd, lda};
    p.B = {b_hi, b_lo, N, Kd, ldb};
    p.C = C; p.c_rows = M; p.c_cols = N; p.ldc = ldc;
    p.M = M; p.N = N; p.Kd = Kd;
    p.negate = flags & 1; p.accumulate = (flags >> 1) & 1; p.lower_tiles_only = (flags >> 2) & 1;
    p.a_tri = (flags >> 4) & 3; p.b_tri = (flags >> 6) & 3;
    return tgemm::launch(p, (cudaStream_t)stream);
}

// AWQ reconstruction loss of a single-Linear parent in Gram form (SURVEY.md B.3, section 7 hard part 6):
//     *loss += sum_{m,n} (D G)[m][n] * E[m][n] = tr(D G D^T) = || X D^T ||_F^2     with G = X^T X (symmetric),
// D = candidate weight - weight [M, K] as bf16 (the MMA operand), E the same matrix in fp32, G [K, K] bf16.  One
// kind::f16 tcgen05 GEMM (fp32 accumulation in TMEM) whose epilogue multiplies the accumulator tile by the E tile
// and reduces: neither the candidate output [T, M] nor D G is ever written.  K % 64 == 0; 16-byte aligned.
int qt_awq_gram_loss(const void* D_bf16, const float* E, const void* G_bf16, int M, int K, double* loss, void* stream) {
    if (!D_bf16 || !E || !G_bf16 || !loss || M <= 0 || K <= 0 || (K % 64)) return QT_ERR_INVALID;
    if (((uintptr_t)D_bf16 | (uintptr_t)G_bf16 | (uintptr_t)E) & 15) return QT_ERR_INVALID;
    tgemm::Problem p;
    p.A = {reinterpret_cast<const float*>(D_bf16), nullptr, M, K, K};
    p.B = {reinterpret_cast<const float*>(G_bf16), nullptr, K, K, K};
    p.C = nullptr; p.c_rows = M; p.c_cols = K; p.ldc = K;
    p.M = M; p.N = K; p.Kd = K;
    p.single = true; p.bf16 = true;
    p.E = E; p.lde = K; p.loss = loss;
    return tgemm::launch(p, (cudaStream_t)stream);
}

}  // extern "C"
