// GPTQ lazy-batch update on the tensor cores:  W[:, i2:] -= Err[:, 0:128] * U[i1:i1+128, i2:]
//
// Replaces the `W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])` line of UPSTREAM llmcompressor
// gptq_quantize.py `quantize_weight`.  Upstream runs it as an fp32
// matmul; here it is a tcgen05 kind::tf32 GEMM made fp32-faithful by operand splitting
// ("3xTF32"): x ~= hi + lo with hi = tf32(x), lo = tf32(x - hi) (about 22 mantissa bits), and
//     A*B ~= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (dropped term ~2^-22 relative)
// accumulated in fp32 in TMEM, so the result matches an fp32 FFMA GEMM to ~1e-6 relative.
//
// Shape of the work: inner dimension is only 128, so the GEMM is HBM-bound on the C traffic
// (32 FLOP per byte): the point of the tensor cores is to get OUT of the way of the memory
// system.  C is never loaded into the SM: the epilogue writes the (negated, via the
// instruction descriptor's negate-A bit) accumulator tile to shared memory and the TMA engine
// adds it into W in L2 (cp.reduce.async.bulk.tensor .add).
//
// Layouts: A = Err_hi|Err_lo  [M,128] fp32 row-major (K-major operand, SW128: 32 floats per row)
//          B = (U^T)_hi|_lo   [K,K]  fp32 row-major: row = output column, k contiguous (K-major too;
//                             an MN-major tf32 B read straight from U needs the 128B_BASE32B swizzle
//                             and returned zeros with plain SWIZZLE_128B, so U is transposed while
//                             it is split - one pass that is needed anyway)
//          C = W              [M,K]  fp32 row-major, cols i2..
// Tile 128 x 256, Kd = 128 as 4 chunks of 32; stage = A_hi,A_lo (2 x 16 KB) + B_hi,B_lo (2 x 32 KB)
// = 96 KB, 2 stages; 2 TMEM accumulators; 4 epilogue warps each own 32 rows and store 32 x 32
// sub-tiles through their own double-buffered 4 KB staging + TMA reduce.
#include "tc_ptx.cuh"

namespace qt {
namespace lazy {
using namespace qt::tc;

constexpr int BM = 128, BN = 256, KD = 128, KC = 32;   // KC floats = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;                               // tf32
constexpr int STAGES = 2;
constexpr int A_BYTES = BM * KC * 4;                    // 16 KB
constexpr int B_BYTES = BN * KC * 4;                    // 32 KB: 256 rows (output columns) x 32 k
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 96 KB
constexpr int EPI_BYTES = 4 * 2 * 4096;                 // 4 warps x 2 buffers x (32 x 32 fp32)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
constexpr int NTHREADS = 192;

// kind::tf32: D=f32 (bit 4), A=B=TF32 (2 at bits 7,10), negate A (bit 13), A and B both K-major
constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(BN >> 3) << 17) |
           ((uint32_t)(BM >> 4) << 24);
}

struct Args {
    int m_tiles, n_tiles;
    int i1, i2;          // B row offset, B/C column offset
};

__global__ void __launch_bounds__(NTHREADS, 1)
lazy_gemm_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                 const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                 const __grid_constant__ CUtensorMap map_c, Args a) {
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
    const int ntiles = a.m_tiles * a.n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ahi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
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
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                // n-major order: consecutive CTAs share the same B columns (L2 reuse of U rows)
                const int tn = t / a.m_tiles, tm = t - tn * a.m_tiles;
                for (int kc = 0; kc < KD / KC; kc++, it++) {
                    const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty[stage], ph ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    uint8_t* s = smem + stage * STAGE_BYTES;
                    tma_load_2d(s, &map_ahi, &full[stage], kc * KC, tm * BM);
                    tma_load_2d(s + A_BYTES, &map_alo, &full[stage], kc * KC, tm * BM);
                    uint8_t* sb = s + 2 * A_BYTES;
                    // B operand = U^T (k contiguous): rows = output columns i2 + tn*256 .., 32 k per chunk
                    tma_load_2d(sb, &map_bhi, &full[stage], a.i1 + kc * KC, a.i2 + tn * BN);
                    tma_load_2d(sb + B_BYTES, &map_blo, &full[stage], a.i1 + kc * KC, a.i2 + tn * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc();
            uint32_t it = 0, li = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, li++) {
                const uint32_t acc = li & 1, aph = (li >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kc = 0; kc < KD / KC; kc++, it++) {
                    const uint32_t stage = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full[stage], ph);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + A_BYTES;
                    const uint32_t sb_hi = sa_hi + 2 * A_BYTES;
                    const uint32_t sb_lo = sb_hi + B_BYTES;
#pragma unroll
                    for (int k = 0; k < KC / UMMA_K; k++) {
                        // A: K-major SW128, rows of 128 B, 8-row groups 1 KB apart; k-step = 32 B inside the row
                        const uint64_t ahi = make_desc_sw128(sa_hi + k * UMMA_K * 4, 16, 1024);
                        const uint64_t alo = make_desc_sw128(sa_lo + k * UMMA_K * 4, 16, 1024);
                        // B: K-major SW128 as well (rows = n)
                        const uint64_t bhi = make_desc_sw128(sb_hi + k * UMMA_K * 4, 16, 1024);
                        const uint64_t blo = make_desc_sw128(sb_lo + k * UMMA_K * 4, 16, 1024);
                        tc_mma_tf32(d_tmem, ahi, bhi, idesc, (kc > 0 || k > 0) ? 1u : 0u);
                        tc_mma_tf32(d_tmem, ahi, blo, idesc, 1u);
                        tc_mma_tf32(d_tmem, alo, bhi, idesc, 1u);
                    }
                    tc_commit(&empty[stage]);
                }
                tc_commit(&tfull[acc]);
            }
        }
    } else {
        const int lg = warp & 3;
        uint8_t* my = epi + (warp - 2) * 2 * 4096;
        uint32_t li = 0, nstore = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, li++) {
            const int tn = t / a.m_tiles, tm = t - tn * a.m_tiles;
            const uint32_t acc = li & 1, aph = (li >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; cc++, nstore++) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + acc * BN + cc * 32, r);
                uint8_t* buf = my + (nstore & 1) * 4096;
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
                    tma_reduce_add_2d(&map_c, buf, a.i2 + tn * BN + cc * 32, tm * BM + lg * 32);
                    bulk_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
        }
        if (lane == 0) bulk_wait<0>();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

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

int launch(const float* err_hi, const float* err_lo, const float* u_hi, const float* u_lo, float* W, int M, int K,
           int i1, int i2, cudaStream_t st) {
    const int ncols = K - i2;
    if (ncols <= 0 || M <= 0) return QT_OK;
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo, mc;
    const CUtensorMapDataType F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    bool ok = make_map_2d(&ma_hi, F32, err_hi, KD, M, KD * 4, KC, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map_2d(&ma_lo, F32, err_lo, KD, M, KD * 4, KC, BM, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map_2d(&mb_hi, F32, u_hi, K, K, (uint64_t)K * 4, KC, BN, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map_2d(&mb_lo, F32, u_lo, K, K, (uint64_t)K * 4, KC, BN, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_map_2d(&mc, F32, W, K, M, (uint64_t)K * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { set_last_error("lazy_gemm tensor maps", cudaErrorInvalidValue); return QT_ERR_CUDA; }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(lazy_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) { set_last_error("lazy_gemm smem attr", e); return QT_ERR_CUDA; }
        attr_set = true;
    }
    Args a;
    a.m_tiles = (M + BM - 1) / BM;
    a.n_tiles = (ncols + BN - 1) / BN;
    a.i1 = i1;
    a.i2 = i2;
    int dev = 0, nsm = kNumSMs;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = a.m_tiles * a.n_tiles;
    lazy_gemm_kernel<<<ntiles < nsm ? ntiles : nsm, NTHREADS, SMEM_BYTES, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, mc, a);
    return check_launch("lazy_gemm");
}

}  // namespace lazy
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
    lazy::split_tf32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, hi, lo, n4);
    return check_launch("split_tf32");
}

// ut_hi/ut_lo [K,K] = transposed tf32 split of U [K,K] (B operand of the lazy update)
int qt_split_tf32_transpose(const float* U, float* ut_hi, float* ut_lo, int K, void* stream) {
    if (!U || !ut_hi || !ut_lo || K <= 0) return QT_ERR_INVALID;
    dim3 grid((K + 31) / 32, (K + 31) / 32);
    lazy::split_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(U, ut_hi, ut_lo, K);
    return check_launch("split_tf32_transpose");
}

// W[:, i2:] -= Err * U[i1:i1+128, i2:] with Err = err_hi + err_lo ([M,128] each) and U^T = u_hi + u_lo ([K,K] each,
// from qt_split_tf32_transpose).
// K % 4 == 0; all pointers 16-byte aligned.
int qt_gptq_lazy_update_tf32x3(const float* err_hi, const float* err_lo, const float* u_hi, const float* u_lo,
                               float* W, int M, int K, int i1, int i2, void* stream) {
    if (!err_hi || !err_lo || !u_hi || !u_lo || !W || M <= 0 || K <= 0 || (K & 3) || i1 < 0 || i2 < 0 || i2 > K)
        return QT_ERR_INVALID;
    if (i1 + 128 > K) return QT_ERR_INVALID;
    return lazy::launch(err_hi, err_lo, u_hi, u_lo, W, M, K, i1, i2, (cudaStream_t)stream);
}

}  // extern "C"
