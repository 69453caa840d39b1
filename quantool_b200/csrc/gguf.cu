// GGUF block packers / unpackers for sm_100a.  Compiled with --fmad=false: every product
// and sum below is a separately rounded fp32 op, in llama.cpp's source order, so the packed
// bytes are bit-identical to the strict-IEEE oracle (oracle/ggml_quants.c).
//
// Replaces: the `llama-quantize` child process of the reference,
//   ref/src/quantool/methods/llama_cpp/llama_cpp.py:165-178 (`GGUF._quantize_gguf`)
// i.e. llama.cpp ggml-quants.c quantize_row_{q8_0,q4_0,q4_1,q5_0,q5_1,q2_K,q3_K,q4_K,q5_K,q6_K}_ref, quantize_iq4_nl
// and dequantize_row_* (§D.1-§D.5).
//
// Data layout in HBM: src is the tensor as a flat row-major array (rows are a multiple of the
// block size, so blocks never straddle rows and the whole tensor is one array of blocks);
// dst is the array of packed blocks, back to back, no padding (GGUF tensor data layout).
//
// Simple types (32-element blocks) are HBM-bound streaming kernels:
//   4 lanes per block, each owning the element pairs that share an output byte, 4 blocks in
//   flight per thread, cross-lane min/max by shuffles, packed bytes staged in shared memory
//   and written back with coalesced 128-bit stores.
// K-quants (256-element super-blocks) run a 19-21 candidate scale search per sub-block with
//   strictly ordered fp32 sums: ~500 ALU ops per element, so they are fp32-ALU-bound, not
//   HBM-bound; one thread owns one sub-block, input/output staged through shared memory.
#include <vector>

#include "gguf_kquant.cuh"

namespace qt {
namespace gguf {

enum : int {
    T_Q4_0 = 2, T_Q4_1 = 3, T_Q5_0 = 6, T_Q5_1 = 7, T_Q8_0 = 8,
    T_Q2_K = 10, T_Q3_K = 11, T_Q4_K = 12, T_Q5_K = 13, T_Q6_K = 14, T_IQ4_NL = 20
};

__host__ __device__ constexpr int block_elems(int t) {
    return (t == T_Q2_K || t == T_Q3_K || t == T_Q4_K || t == T_Q5_K || t == T_Q6_K) ? 256
         : (t == T_Q4_0 || t == T_Q4_1 || t == T_Q5_0 || t == T_Q5_1 || t == T_Q8_0 || t == T_IQ4_NL) ? 32 : -1;
}
__host__ __device__ constexpr int block_bytes(int t) {
    return (t == T_Q4_0 || t == T_IQ4_NL) ? 18 : t == T_Q4_1 ? 20 : t == T_Q5_0 ? 22 : t == T_Q5_1 ? 24 : t == T_Q8_0 ? 34
         : t == T_Q2_K ? 84 : t == T_Q3_K ? 110 : t == T_Q4_K ? 144 : t == T_Q5_K ? 176 : t == T_Q6_K ? 210 : -1;
}

QT_D void sts16(uint8_t* p, uint32_t v) { *reinterpret_cast<unsigned short*>(p) = (unsigned short)v; }

// spread the 4 nibbles of a 16-bit value into the low nibbles of 4 bytes
QT_D uint32_t spread4(uint32_t v) {
    return (v & 0xFu) | ((v & 0xF0u) << 4) | ((v & 0xF00u) << 8) | ((v & 0xF000u) << 12);
}

// copy `bytes` from shared staging to global with 128-bit stores (+ byte tail on the last CTA)
QT_D void copy_out(uint8_t* __restrict__ gdst, const uint8_t* sout, int bytes) {
    const int nvec = bytes >> 4;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x)
        stg_stream(gdst + 16 * i, *reinterpret_cast<const uint4*>(sout + 16 * i));
    for (int i = (nvec << 4) + threadIdx.x; i < bytes; i += blockDim.x) gdst[i] = sout[i];
}

// ---------------------------------------------------------------------------------------
// 32-element block types.  A warp covers 8 blocks per pass (4 lanes each), U passes per iteration.
// Lane q (0..3) of a block owns elements 4q..4q+3 and 16+4q..16+4q+3: exactly the pairs that
// share an output byte in the 4/5-bit formats, so no packed data crosses lanes.  The kernels
// were issue-bound (29 instructions per element in the first version): the arg-max is tracked
// as a signed max and min (2 FMNMX per element; the first-wins tie between +a and -a is a rare
// slow path), and float->int truncation of the non-negative codes uses a round-down magic add
// instead of the quarter-rate F2I.
// ---------------------------------------------------------------------------------------
// One launch can pack many tensors of the same type ("segments"): a model is hundreds of small tensors and a
// launch per tensor made BASELINE config 1 (SmolLM2-135M, 211 tensors) launch-bound (3.2 ms for 0.4 GB).
// A tile = the work of one CTA iteration; tile_start[] is the exclusive prefix sum of tiles per segment.
struct Seg { const void* src; uint8_t* dst; long long n; };   // n = blocks (32-element types) or super-blocks
struct Segs {
    int nseg, total_tiles;
    const int* tile_start;   // device [nseg + 1]; unused when nseg == 1
    const Seg* table;        // device [nseg];     unused when nseg == 1
    Seg one;
};
QT_D Seg seg_of(const Segs& s, int tile, int& local_tile) {
    if (s.nseg == 1) { local_tile = tile; return s.one; }
    int lo = 0, hi = s.nseg - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(s.tile_start + mid) <= tile) lo = mid; else hi = mid - 1;
    }
    local_tile = tile - __ldg(s.tile_start + lo);
    return s.table[lo];
}

constexpr int kSimpleU = 4;
constexpr int kSimpleBlocksPerCta = 64 * kSimpleU;  // 256 blocks: 256*BB is a multiple of 16

QT_D uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// elements 4q..4q+3 -> v[0..3], 16+4q..16+4q+3 -> v[4..7] of block `blk`
template <int DT>
QT_D void load_block_quarter(const void* __restrict__ src, int64_t blk, int q, float v[8]) {
    if (DT == QT_F32) {
        const char* p = (const char*)src + blk * 128 + q * 16;
        const uint4 a = ldg_stream(p), b = ldg_stream(p + 64);
        v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
        v[4] = __uint_as_float(b.x); v[5] = __uint_as_float(b.y); v[6] = __uint_as_float(b.z); v[7] = __uint_as_float(b.w);
    } else {
        const char* p = (const char*)src + blk * 64 + q * 8;
        const uint2 a = ldg_stream8(p), b = ldg_stream8(p + 32);
        const uint32_t w[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (DT == QT_F16) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                v[2 * i] = f.x; v[2 * i + 1] = f.y;
            } else {
                v[2 * i] = __uint_as_float(w[i] << 16);
                v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            }
        }
    }
}

// low bytes of four 32-bit values -> one word (3 PRMT)
QT_D uint32_t gather_b0(float a, float b, float c, float d) {
    const uint32_t ab = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x0040);
    const uint32_t cd = __byte_perm(__float_as_uint(c), __float_as_uint(d), 0x0040);
    return __byte_perm(ab, cd, 0x5410);
}

// Float -> code without F2I: a round-down add of 2^23 to t in [0, 2^22) leaves floor(t) in the low
// mantissa bits.  (C's (int8_t)(x + 8.5f) truncates; the argument is >= 0.)
constexpr float kMagic = 8388608.0f;

template <int TYPE>
QT_D void pack_simple_block(const float (&v)[8], int q, uint8_t* o) {
    const unsigned m4 = 0xFu << (threadIdx.x & 28);   // the 4 lanes of this block
    if (TYPE == T_Q8_0) {
        // Q8_0 keeps the plain mapping: lane q owns elements 8q..8q+7 (one 128-bit load)
        float amax = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) amax = fmaxf(amax, fabsf(v[i]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        const float d = amax / 127;
        const float id = d ? 1.0f / d : 0.0f;
        uint32_t w[2] = {0, 0};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int qi = (int)roundf(v[i] * id);
            w[i >> 2] |= (uint32_t)(qi & 0xff) << (8 * (i & 3));
        }
        if (q == 0) sts16(o, __half_as_ushort(__float2half_rn(d)));
        uint8_t* qs = o + 2 + 8 * q;
        sts16(qs, w[0] & 0xffff); sts16(qs + 2, w[0] >> 16);
        sts16(qs + 4, w[1] & 0xffff); sts16(qs + 6, w[1] >> 16);
        return;
    }
    constexpr bool kSym = (TYPE == T_Q4_0 || TYPE == T_Q5_0);
    constexpr bool k5 = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
    constexpr int kHdr = kSym ? 2 : 4;
    float mn = v[0], mx = v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) { mn = fminf(mn, v[i]); mx = fmaxf(mx, v[i]); }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    // Codes stay in the low mantissa byte of "magic" floats (2^23 + n, see kMagic): the byte
    // gathers below read only that byte, so no mask or F2I is needed.
    float tf[8];
    if (kSym) {
        // signed value of the first element with the largest |v| (strict < in a forward scan)
        float smax = (mx > -mn) ? mx : mn;
        if (mx == -mn && mx > 0.f) {
            // +a and -a both present: the earlier element decides the sign.  Block-uniform, rare.
            int key = 64;
#pragma unroll
            for (int i = 7; i >= 0; i--) {
                const int e = (i < 4) ? 4 * q + i : 16 + 4 * q + (i - 4);
                if (fabsf(v[i]) == mx) { const int k2 = 2 * e + (v[i] < 0.f ? 1 : 0); key = k2 < key ? k2 : key; }
            }
            int o1 = __shfl_xor_sync(m4, key, 1); key = o1 < key ? o1 : key;
            o1 = __shfl_xor_sync(m4, key, 2); key = o1 < key ? o1 : key;
            smax = (key & 1) ? mn : mx;
        }
        const float d = smax / (k5 ? -16 : -8);
        const float id = d ? 1.0f / d : 0.0f;
        if (q == 0) sts16(o, __half_as_ushort(__float2half_rn(d)));
        // MIN(15, (int8_t)(x*id + 8.5f)): the clamp commutes with the truncation
#pragma unroll
        for (int i = 0; i < 8; i++)
            tf[i] = __fadd_rd(fminf(v[i] * id + (k5 ? 16.5f : 8.5f), k5 ? 31.0f : 15.0f), kMagic);
    } else {
        const float d = (mx - mn) / (k5 ? 31 : 15);
        const float id = d ? 1.0f / d : 0.0f;
        if (q == 0) {
            sts16(o, __half_as_ushort(__float2half_rn(d)));
            sts16(o + 2, __half_as_ushort(__float2half_rn(mn)));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float t = (v[i] - mn) * id + 0.5f;
            tf[i] = __fadd_rd(k5 ? t : fminf(t, 15.0f), kMagic);   // Q5_1 casts to uint8_t unclamped
        }
    }
    // lo = codes of elements 4q..4q+3, hi = codes of 16+4q..16+4q+3, one byte each
    const uint32_t lo = gather_b0(tf[0], tf[1], tf[2], tf[3]);
    const uint32_t hi = gather_b0(tf[4], tf[5], tf[6], tf[7]);
    // byte j = elem j (low nibble) | elem j+16 (high nibble), j = 4q + i: both live in this lane
    uint32_t word;
    if (k5) {
        word = (lo & 0x0F0F0F0Fu) | ((hi << 4) & 0xF0F0F0F0u);
        // bit 4 of each of the four bytes -> a nibble (the multiply lines them up at bits 28..31)
        const uint32_t nl = ((lo & 0x10101010u) * 0x01020408u) >> 28;
        const uint32_t nh = ((hi & 0x10101010u) * 0x01020408u) >> 28;
        uint32_t qh = (nl << (4 * q)) | (nh << (16 + 4 * q));
        qh |= __shfl_xor_sync(0xffffffffu, qh, 1);
        qh |= __shfl_xor_sync(0xffffffffu, qh, 2);
        if (q == 0) { sts16(o + kHdr, qh & 0xffff); sts16(o + kHdr + 2, qh >> 16); }
    } else {
        word = lo + (hi << 4);   // codes are <= 15
    }
    uint8_t* qs = o + kHdr + (k5 ? 4 : 0) + 4 * q;
    sts16(qs, word & 0xffff);
    sts16(qs + 2, word >> 16);
}

template <int TYPE, int DT, bool VIA_F16>
__global__ void __launch_bounds__(256) pack_simple_kernel(const Segs segs) {
    // Each warp owns 32 consecutive blocks per iteration (8 blocks x kSimpleU passes): 32*BB bytes
    // of output is a multiple of 16 for every type, so the warp stages and stores its own slice
    // and the CTA never synchronises - warps drift apart and cover each other's load latency.
    constexpr int BB = block_bytes(TYPE);
    constexpr int kWarpBlocks = 8 * kSimpleU;
    static_assert((kWarpBlocks * BB) % 16 == 0, "warp slice must be 16-byte granular");
    __shared__ __align__(16) uint8_t sout[kSimpleBlocksPerCta * BB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3, bl = lane >> 2;
    uint8_t* wout = sout + warp * kWarpBlocks * BB;
    for (int tile = blockIdx.x; tile < segs.total_tiles; tile += gridDim.x) {
        int lt;
        const Seg sg = seg_of(segs, tile, lt);
        const void* __restrict__ src = sg.src;
        uint8_t* __restrict__ dst = sg.dst;
        const int64_t nblocks = sg.n;
        const int64_t base = (int64_t)lt * kSimpleBlocksPerCta + warp * kWarpBlocks;
        if (base >= nblocks) continue;     // warp-uniform; this kernel has no CTA-wide barrier
        float v[kSimpleU][8];
        const bool full = base + kWarpBlocks <= nblocks;
        if (full) {
            // one base pointer, compile-time offsets, all loads issued before any math
#pragma unroll
            for (int u = 0; u < kSimpleU; u++) {
                if (TYPE == T_Q8_0) load8<DT>(src, (base + bl) * 4 + q + u * 32, v[u]);
                else load_block_quarter<DT>(src, base + bl + u * 8, q, v[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kSimpleU; u++) {
                const int64_t blk = base + u * 8 + bl;
                if (blk < nblocks) {
                    if (TYPE == T_Q8_0) load8<DT>(src, blk * 4 + q, v[u]);
                    else load_block_quarter<DT>(src, blk, q, v[u]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[u][i] = 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kSimpleU; u++) {
            if (VIA_F16 && DT != QT_F16) {
#pragma unroll
                for (int i = 0; i < 8; i++) v[u][i] = round_via_f16(v[u][i]);
            }
            pack_simple_block<TYPE>(v[u], q, wout + (u * 8 + bl) * BB);
        }
        __syncwarp();
        uint8_t* gdst = dst + base * BB;
        if (full) {
#pragma unroll
            for (int i = lane; i < kWarpBlocks * BB / 16; i += 32)
                stg_stream(gdst + 16 * i, *reinterpret_cast<const uint4*>(wout + 16 * i));
        } else {
            const int bytes = (int)(nblocks - base) * BB;
            for (int i = lane; i < (bytes >> 4); i += 32)
                stg_stream(gdst + 16 * i, *reinterpret_cast<const uint4*>(wout + 16 * i));
            for (int i = (bytes & ~15) + lane; i < bytes; i += 32) gdst[i] = wout[i];
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------
// K-quants
// ---------------------------------------------------------------------------------------
// stage NELEM consecutive elements (from super-block `sb0`) into s.x rows of ROW elements
template <int DT, bool VIA_F16, int ROW, int PAD, int NELEM>
QT_D void stage_in(const void* __restrict__ src, int64_t elem0, int64_t nvalid_elems, float (*sx)[PAD]) {
    for (int c = threadIdx.x; c < NELEM / 8; c += blockDim.x) {
        float v[8];
        if ((int64_t)c * 8 < nvalid_elems) {
            load8<DT>(src, elem0 / 8 + c, v);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.f;
        }
        const int e = c * 8;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float f = v[i];
            if (VIA_F16 && DT != QT_F16) f = round_via_f16(f);
            sx[(e + i) / ROW][(e + i) % ROW] = f;
        }
    }
}

constexpr int kK45Nsb = 32;  // 256 threads, 32 super-blocks (8192 elements) per CTA iteration
constexpr int kK6Nsb = 16;   // 256 threads, 16 super-blocks (4096 elements)

template <int TYPE, int DT, bool VIA_F16>
__global__ void __launch_bounds__(256) pack_k45_kernel(const Segs segs) {
    constexpr int BB = block_bytes(TYPE);
    using S = kq::K45Shared<kK45Nsb, BB>;
    __shared__ S s;
    const int t = threadIdx.x;
    constexpr int nmax = TYPE == T_Q4_K ? 15 : 31;
    for (int tile = blockIdx.x; tile < segs.total_tiles; tile += gridDim.x) {
        int lt;
        const Seg sg = seg_of(segs, tile, lt);
        const void* __restrict__ src = sg.src;
        uint8_t* __restrict__ dst = sg.dst;
        const int64_t base = (int64_t)lt * kK45Nsb;
        const int64_t left = sg.n - base;
        const int nvalid = left < kK45Nsb ? (int)left : kK45Nsb;
        stage_in<DT, VIA_F16, 32, 33, kK45Nsb * 256>(src, base * 256, (int64_t)nvalid * 256, s.u.x);
        __syncthreads();
        kq::K45Thread th;
        if (TYPE == T_Q4_K) kq::k45_phase_a(t, s, th, 15, -1.f, 0.1f, 20);
        else                kq::k45_phase_a(t, s, th, 31, -0.5f, 0.1f, 15);
        __syncthreads();
        kq::k45_phase_b<S, BB>(t, s, th, nmax);
        __syncthreads();
        if (TYPE == T_Q4_K) kq::q4k_phase_c(t, s); else kq::q5k_phase_c(t, s);
        __syncthreads();
        copy_out(dst + base * BB, s.u.o.out, nvalid * BB);
        __syncthreads();
    }
}

template <int DT, bool VIA_F16>
__global__ void __launch_bounds__(256) pack_q6k_kernel(const Segs segs) {
    using S = kq::K6Shared<kK6Nsb>;
    __shared__ S s;
    const int t = threadIdx.x;
    for (int tile = blockIdx.x; tile < segs.total_tiles; tile += gridDim.x) {
        int lt;
        const Seg sg = seg_of(segs, tile, lt);
        const void* __restrict__ src = sg.src;
        uint8_t* __restrict__ dst = sg.dst;
        const int64_t base = (int64_t)lt * kK6Nsb;
        const int64_t left = sg.n - base;
        const int nvalid = left < kK6Nsb ? (int)left : kK6Nsb;
        stage_in<DT, VIA_F16, 16, 17, kK6Nsb * 256>(src, base * 256, (int64_t)nvalid * 256, s.x);
        __syncthreads();
        kq::K6Thread th;
        kq::q6k_phase_a(t, s, th);
        __syncthreads();
        kq::q6k_phase_b(t, s, th);
        __syncthreads();
        kq::q6k_phase_c(t, s);
        __syncthreads();
        copy_out(dst + base * 210, s.out, nvalid * 210);
        __syncthreads();
    }
}

template <int TYPE, int DT, bool VIA_F16>
__global__ void __launch_bounds__(256) pack_k23_kernel(const Segs segs) {
    constexpr int BB = block_bytes(TYPE);
    using S = kq::K23Shared<kK6Nsb, BB>;
    __shared__ S s;
    const int t = threadIdx.x;
    for (int tile = blockIdx.x; tile < segs.total_tiles; tile += gridDim.x) {
        int lt;
        const Seg sg = seg_of(segs, tile, lt);
        const void* __restrict__ src = sg.src;
        uint8_t* __restrict__ dst = sg.dst;
        const int64_t base = (int64_t)lt * kK6Nsb;
        const int64_t left = sg.n - base;
        const int nvalid = left < kK6Nsb ? (int)left : kK6Nsb;
        stage_in<DT, VIA_F16, 16, 17, kK6Nsb * 256>(src, base * 256, (int64_t)nvalid * 256, s.x);
        __syncthreads();
        if constexpr (TYPE == T_Q2_K) {
            kq::K2Thread th;
            kq::q2k_phase_a(t, s, th);
            __syncthreads();
            kq::q2k_phase_b(t, s, th);
            __syncthreads();
            kq::q2k_phase_c(t, s);
        } else {
            kq::K3Thread th;
            kq::q3k_phase_a(t, s, th);
            __syncthreads();
            kq::q3k_phase_b(t, s, th);
            __syncthreads();
            kq::q3k_phase_c(t, s);
        }
        __syncthreads();
        copy_out(dst + base * BB, s.out, nvalid * BB);
        __syncthreads();
    }
}

// IQ4_NL: one thread per 32-element block (16 candidate scales x 32 table look-ups: ALU-bound like the K-quants)
template <int DT, bool VIA_F16>
__global__ void __launch_bounds__(256) pack_iq4nl_kernel(const Segs segs) {
    __shared__ float sx[kSimpleBlocksPerCta][33];
    __shared__ __align__(16) uint8_t sout[kSimpleBlocksPerCta * 18];
    __shared__ float tbl[16];
    const int t = threadIdx.x;
    if (t < 16) {
        const float vals[16] = QT_IQ4NL_VALUES;
        tbl[t] = vals[t];
    }
    for (int tile = blockIdx.x; tile < segs.total_tiles; tile += gridDim.x) {
        int lt;
        const Seg sg = seg_of(segs, tile, lt);
        const void* __restrict__ src = sg.src;
        uint8_t* __restrict__ dst = sg.dst;
        const int64_t base = (int64_t)lt * kSimpleBlocksPerCta;
        const int64_t left = sg.n - base;
        const int nvalid = left < kSimpleBlocksPerCta ? (int)left : kSimpleBlocksPerCta;
        stage_in<DT, VIA_F16, 32, 33, kSimpleBlocksPerCta * 32>(src, base * 32, (int64_t)nvalid * 32, sx);
        __syncthreads();
        float v[32];
#pragma unroll
        for (int l = 0; l < 32; ++l) v[l] = sx[t][l];
        kq::iq4nl_block(tbl, v, sout + t * 18);
        __syncthreads();
        copy_out(dst + base * 18, sout, nvalid * 18);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// dequantize: one thread per 16 (simple) / 32 (K) output elements; packed blocks are read
// byte-wise through the read-only path (they are 2-byte aligned at best), outputs are
// written as 128-bit stores.  Expressions follow llama.cpp dequantize_row_* / gguf-py.
// ---------------------------------------------------------------------------------------
QT_D float ld_f16(const uint8_t* p) {
    return __half2float(__ushort_as_half((unsigned short)(p[0] | (p[1] << 8))));
}
QT_D void st_f4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

template <int TYPE>
__global__ void __launch_bounds__(256) dequant_simple_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                             int64_t nblocks) {
    constexpr int BB = block_bytes(TYPE);
    // two threads per block: half h covers output elements 16h..16h+15
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t blk = gid >> 1;
    const int h = (int)(gid & 1);
    if (blk >= nblocks) return;
    const uint8_t* b = src + blk * BB;
    float* y = dst + blk * 32 + 16 * h;
    float out[16];
    if (TYPE == T_Q8_0) {
        const float d = ld_f16(b);
#pragma unroll
        for (int j = 0; j < 16; j++) out[j] = (float)(int)(signed char)b[2 + 16 * h + j] * d;
    } else if (TYPE == T_IQ4_NL) {
        const float d = ld_f16(b);
#pragma unroll
        for (int j = 0; j < 16; j++) out[j] = d * kq::iq4nl_value(h ? (b[2 + j] >> 4) : (b[2 + j] & 0xF));
    } else {
        constexpr bool kSym = (TYPE == T_Q4_0 || TYPE == T_Q5_0);
        constexpr bool k5 = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
        constexpr int kHdr = kSym ? 2 : 4;
        const float d = ld_f16(b);
        const float m = kSym ? 0.f : ld_f16(b + 2);
        uint32_t qh = 0;
        if (k5) qh = b[kHdr] | (b[kHdr + 1] << 8) | (b[kHdr + 2] << 16) | ((uint32_t)b[kHdr + 3] << 24);
        const uint8_t* qs = b + kHdr + (k5 ? 4 : 0);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            int x = h ? (qs[j] >> 4) : (qs[j] & 0xF);
            if (k5) x |= (int)((qh >> (j + 16 * h)) & 1) << 4;
            if (kSym) out[j] = (float)(x - (k5 ? 16 : 8)) * d;
            else out[j] = (float)x * d + m;
        }
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4) st_f4(y + j, out[j], out[j + 1], out[j + 2], out[j + 3]);
}

QT_D void scale_min_k4(int j, const uint8_t* q, int& d, int& m) {
    if (j < 4) { d = q[j] & 63; m = q[j + 4] & 63; }
    else { d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

template <int TYPE>
__global__ void __launch_bounds__(256) dequant_k_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                        int64_t nsuper) {
    constexpr int BB = block_bytes(TYPE);
    // 8 threads per super-block, thread j covers output elements 32j..32j+31
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t sb = gid >> 3;
    const int j = (int)(gid & 7);
    if (sb >= nsuper) return;
    const uint8_t* b = src + sb * BB;
    float* y = dst + sb * 256 + 32 * j;
    float out[32];
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        const float d = ld_f16(b), mn = ld_f16(b + 2);
        int sc, m;
        scale_min_k4(j, b + 4, sc, m);
        const float d1 = d * sc, m1 = mn * m;
        const int c = j >> 1, hi = j & 1;
        const uint8_t* q = b + (TYPE == T_Q4_K ? 16 : 48) + 32 * c;
        const uint8_t* qh = b + 16;
#pragma unroll
        for (int l = 0; l < 32; l++) {
            int x = hi ? (q[l] >> 4) : (q[l] & 0xF);
            if (TYPE == T_Q5_K) x += ((qh[l] >> j) & 1) ? 16 : 0;
            out[l] = d1 * x - m1;
        }
    } else if (TYPE == T_Q2_K || TYPE == T_Q3_K) {
        // element e = 32j + l -> half n = j/4, shift 2*(j%4); sub-block scale index 2j + l/16
        const int n = j >> 2, sh = 2 * (j & 3);
        const uint8_t* q = b + (TYPE == T_Q2_K ? 16 : 32) + 32 * n;
        if (TYPE == T_Q2_K) {
            const float d = ld_f16(b + 80), mn = ld_f16(b + 82);
#pragma unroll
            for (int l = 0; l < 32; l++) {
                const uint8_t sc = b[2 * j + l / 16];
                const float dl = d * (sc & 0xF), ml = mn * (sc >> 4);
                out[l] = dl * (float)((q[l] >> sh) & 3) - ml;
            }
        } else {
            const float d = ld_f16(b + 108);
            const uint8_t* scb = b + 96;
#pragma unroll
            for (int l = 0; l < 32; l++) {
                const int is = 2 * j + l / 16;
                const int lo = is < 8 ? (scb[is] & 0xF) : (scb[is - 8] >> 4);
                const int hi = (scb[8 + is % 4] >> (2 * (is / 4))) & 3;
                const float dl = d * (float)((int)(signed char)(lo | (hi << 4)) - 32);
                out[l] = dl * (float)((int)((q[l] >> sh) & 3) - (((b[l] >> j) & 1) ? 0 : 4));
            }
        }
    } else {  // Q6_K: element e = 32j + l -> half n=e/128, quarter k=(e%128)/32
        const float d = ld_f16(b + 208);
        const int n = j >> 2, k = j & 3;
        const uint8_t* ql = b + 64 * n;
        const uint8_t* qh = b + 128 + 32 * n;
        const int8_t* sc = reinterpret_cast<const int8_t*>(b + 192) + 8 * n;
#pragma unroll
        for (int l = 0; l < 32; l++) {
            const int is = l / 16;
            const int lo = (k & 1) ? ql[l + 32] : ql[l];
            const int nibble = (k >= 2) ? (lo >> 4) : (lo & 0xF);
            const int q = (int)(signed char)(nibble | (((qh[l] >> (2 * k)) & 3) << 4)) - 32;
            out[l] = d * sc[is + 2 * k] * q;
        }
    }
#pragma unroll
    for (int l = 0; l < 32; l += 4) st_f4(y + l, out[l], out[l + 1], out[l + 2], out[l + 3]);
}

// ---- staged form: 128-bit loads of the packed bytes, fully coalesced 128-bit stores -------------------------------
// The kernels above read every packed byte with its own global load and store 64-128 contiguous bytes per THREAD
// (a warp-wide store touches 32 half-filled sectors): 25-55 % of HBM peak, at or below vLLM's ggml_dequantize.  Here
// a CTA stages the packed bytes of 8192 output elements in shared memory with 128-bit loads, and thread t of step k
// decodes elements 1024 k + 4 t .. + 3, so that consecutive lanes store consecutive float4s.  One element is decoded by
// `dq_elem` with exactly the expressions of the kernels above (llama.cpp dequantize_row_* / gguf-py).
template <int TYPE>
QT_D float dq_elem(const uint8_t* b, int e) {
    if (TYPE == T_Q8_0) {
        return (float)(int)(signed char)b[2 + e] * ld_f16(b);
    } else if (TYPE == T_IQ4_NL) {
        const uint8_t q = b[2 + (e & 15)];
        return ld_f16(b) * kq::iq4nl_value(e >= 16 ? (q >> 4) : (q & 0xF));
    } else if (TYPE == T_Q4_0 || TYPE == T_Q4_1 || TYPE == T_Q5_0 || TYPE == T_Q5_1) {
        constexpr bool kSym = (TYPE == T_Q4_0 || TYPE == T_Q5_0);
        constexpr bool k5 = (TYPE == T_Q5_0 || TYPE == T_Q5_1);
        constexpr int kHdr = kSym ? 2 : 4;
        const float d = ld_f16(b);
        const float m = kSym ? 0.f : ld_f16(b + 2);
        const uint8_t* qs = b + kHdr + (k5 ? 4 : 0);
        const uint8_t q = qs[e & 15];
        int x = e >= 16 ? (q >> 4) : (q & 0xF);
        if (k5) {
            const uint32_t qh = b[kHdr] | (b[kHdr + 1] << 8) | (b[kHdr + 2] << 16) | ((uint32_t)b[kHdr + 3] << 24);
            x |= (int)((qh >> e) & 1) << 4;
        }
        if (kSym) return (float)(x - (k5 ? 16 : 8)) * d;
        return (float)x * d + m;
    } else if (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        const int j = e >> 5, l = e & 31;
        const float d = ld_f16(b), mn = ld_f16(b + 2);
        int sc, m;
        scale_min_k4(j, b + 4, sc, m);
        const float d1 = d * sc, m1 = mn * m;
        const uint8_t* q = b + (TYPE == T_Q4_K ? 16 : 48) + 32 * (j >> 1);
        int x = (j & 1) ? (q[l] >> 4) : (q[l] & 0xF);
        if (TYPE == T_Q5_K) x += ((b[16 + l] >> j) & 1) ? 16 : 0;
        return d1 * x - m1;
    } else if (TYPE == T_Q2_K) {
        const int j = e >> 5, l = e & 31;
        const int n = j >> 2, sh = 2 * (j & 3);
        const float d = ld_f16(b + 80), mn = ld_f16(b + 82);
        const uint8_t sc = b[2 * j + l / 16];
        const float dl = d * (sc & 0xF), ml = mn * (sc >> 4);
        return dl * (float)((b[16 + 32 * n + l] >> sh) & 3) - ml;
    } else if (TYPE == T_Q3_K) {
        const int j = e >> 5, l = e & 31;
        const int n = j >> 2, sh = 2 * (j & 3);
        const float d = ld_f16(b + 108);
        const uint8_t* scb = b + 96;
        const int is = 2 * j + l / 16;
        const int lo = is < 8 ? (scb[is] & 0xF) : (scb[is - 8] >> 4);
        const int hi = (scb[8 + is % 4] >> (2 * (is / 4))) & 3;
        const float dl = d * (float)((int)(signed char)(lo | (hi << 4)) - 32);
        return dl * (float)((int)((b[32 + 32 * n + l] >> sh) & 3) - (((b[l] >> j) & 1) ? 0 : 4));
    } else {  // Q6_K
        const int j = e >> 5, l = e & 31;
        const float d = ld_f16(b + 208);
        const int n = j >> 2, k = j & 3;
        const uint8_t* ql = b + 64 * n;
        const int8_t* sc = reinterpret_cast<const int8_t*>(b + 192) + 8 * n;
        const int lo = (k & 1) ? ql[l + 32] : ql[l];
        const int nibble = (k >= 2) ? (lo >> 4) : (lo & 0xF);
        const int q = (int)(signed char)(nibble | (((b[128 + 32 * n + l] >> (2 * k)) & 3) << 4)) - 32;
        return d * sc[l / 16 + 2 * k] * q;
    }
}

// four consecutive elements p .. p+3 (p % 4 == 0) of one block: the header / sub-block values are derived once.
// Q4_K / Q5_K blocks are 16-byte multiples, so the four code bytes are one aligned 32-bit shared-memory load.
template <int TYPE>
QT_D void dq4(const uint8_t* b, int p, float (&o)[4]) {
    if (TYPE == T_Q4_K || TYPE == T_Q5_K) {
        const int j = p >> 5, l0 = p & 31;
        const float d = ld_f16(b), mn = ld_f16(b + 2);
        int sc, m;
        scale_min_k4(j, b + 4, sc, m);
        const float d1 = d * sc, m1 = mn * m;
        const uint32_t qq = *reinterpret_cast<const uint32_t*>(b + (TYPE == T_Q4_K ? 16 : 48) + 32 * (j >> 1) + l0);
        const uint32_t hh = TYPE == T_Q5_K ? *reinterpret_cast<const uint32_t*>(b + 16 + l0) : 0u;
        const int sh = (j & 1) ? 4 : 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int x = (int)((qq >> (8 * i + sh)) & 0xFu);
            if (TYPE == T_Q5_K) x += ((hh >> (8 * i + j)) & 1u) ? 16 : 0;
            o[i] = d1 * x - m1;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = dq_elem<TYPE>(b, p + i);
    }
}

template <int TYPE>
__global__ void __launch_bounds__(256) dequant_staged_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                             int64_t nblocks) {
    constexpr int BE = block_elems(TYPE), BB = block_bytes(TYPE);
    constexpr int NBLK = 8192 / BE;                      // blocks per tile: 256 (32-element types) or 32 (K types)
    static_assert((NBLK * BB) % 16 == 0, "a tile of packed blocks must be 16-byte granular");
    __shared__ __align__(16) uint8_t sb[NBLK * BB];
    const int tid = threadIdx.x;
    const int64_t ntiles = (nblocks + NBLK - 1) / NBLK;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * NBLK;
        const int nv = (int)((nblocks - base) < NBLK ? (nblocks - base) : NBLK);
        const int bytes = nv * BB;
        const uint8_t* g = src + base * BB;
        for (int i = tid; i < (bytes >> 4); i += 256) *reinterpret_cast<uint4*>(sb + 16 * i) = ldg_stream(g + 16 * i);
        for (int i = (bytes & ~15) + tid; i < bytes; i += 256) sb[i] = g[i];
        __syncthreads();
        float* out = dst + base * BE;
        const int nelem = nv * BE;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int e = 1024 * k + 4 * tid;
            if (e < nelem) {
                const uint8_t* b = sb + (e / BE) * BB;
                const int p = e % BE;
                float v[4];
                dq4<TYPE>(b, p, v);
                uint4 u;
                u.x = __float_as_uint(v[0]); u.y = __float_as_uint(v[1]); u.z = __float_as_uint(v[2]); u.w = __float_as_uint(v[3]);
                stg_stream(out + e, u);
            }
        }
        __syncthreads();
    }
}

template <int TYPE>
static void launch_dequant_staged(const uint8_t* s, float* dst, int64_t nblk, cudaStream_t st) {
    const int64_t ntiles = (nblk + 8192 / block_elems(TYPE) - 1) / (8192 / block_elems(TYPE));
    const int64_t cap = (int64_t)kNumSMs * 8;
    dequant_staged_kernel<TYPE><<<(unsigned)(ntiles < cap ? ntiles : cap), 256, 0, st>>>(s, dst, nblk);
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
static int grid_for(int64_t units_per_cta_iter_total, int ctas_per_sm) {
    const int64_t cap = (int64_t)kNumSMs * ctas_per_sm;
    return (int)(units_per_cta_iter_total < cap ? (units_per_cta_iter_total > 0 ? units_per_cta_iter_total : 1) : cap);
}

template <int TYPE>
constexpr int tile_units() {
    return block_elems(TYPE) == 32 ? kSimpleBlocksPerCta : (TYPE == T_Q4_K || TYPE == T_Q5_K) ? kK45Nsb : kK6Nsb;
}

template <int TYPE, int DT, bool VIA>
static void launch_pack(const Segs& segs, cudaStream_t st) {
    if constexpr (TYPE == T_IQ4_NL) {
        pack_iq4nl_kernel<DT, VIA><<<grid_for(segs.total_tiles, 4), 256, 0, st>>>(segs);
    } else if constexpr (block_elems(TYPE) == 32) {
        pack_simple_kernel<TYPE, DT, VIA><<<grid_for(segs.total_tiles, 8), 256, 0, st>>>(segs);
    } else if constexpr (TYPE == T_Q6_K) {
        pack_q6k_kernel<DT, VIA><<<grid_for(segs.total_tiles, 4), 256, 0, st>>>(segs);
    } else if constexpr (TYPE == T_Q2_K || TYPE == T_Q3_K) {
        pack_k23_kernel<TYPE, DT, VIA><<<grid_for(segs.total_tiles, 4), 256, 0, st>>>(segs);
    } else {
        pack_k45_kernel<TYPE, DT, VIA><<<grid_for(segs.total_tiles, 4), 256, 0, st>>>(segs);
    }
}

template <int TYPE>
static int dispatch_dt(const Segs& segs, int dt, int via, cudaStream_t st) {
    switch (dt) {
        case QT_F32: via ? launch_pack<TYPE, QT_F32, true>(segs, st) : launch_pack<TYPE, QT_F32, false>(segs, st); break;
        case QT_F16: launch_pack<TYPE, QT_F16, false>(segs, st); break;
        case QT_BF16: via ? launch_pack<TYPE, QT_BF16, true>(segs, st) : launch_pack<TYPE, QT_BF16, false>(segs, st); break;
        default: return QT_ERR_INVALID;
    }
    return check_launch("qt_gguf_quantize");
}

static int tile_units_of(int t) {
    switch (t) {
        case T_Q4_K: case T_Q5_K: return kK45Nsb;
        case T_Q2_K: case T_Q3_K: case T_Q6_K: return kK6Nsb;
        default: return kSimpleBlocksPerCta;
    }
}

static int dispatch_type(int ggml_type, const Segs& segs, int dt, int via, cudaStream_t st) {
    switch (ggml_type) {
        case T_Q4_0: return dispatch_dt<T_Q4_0>(segs, dt, via, st);
        case T_Q4_1: return dispatch_dt<T_Q4_1>(segs, dt, via, st);
        case T_Q5_0: return dispatch_dt<T_Q5_0>(segs, dt, via, st);
        case T_Q5_1: return dispatch_dt<T_Q5_1>(segs, dt, via, st);
        case T_Q8_0: return dispatch_dt<T_Q8_0>(segs, dt, via, st);
        case T_IQ4_NL: return dispatch_dt<T_IQ4_NL>(segs, dt, via, st);
        case T_Q2_K: return dispatch_dt<T_Q2_K>(segs, dt, via, st);
        case T_Q3_K: return dispatch_dt<T_Q3_K>(segs, dt, via, st);
        case T_Q4_K: return dispatch_dt<T_Q4_K>(segs, dt, via, st);
        case T_Q5_K: return dispatch_dt<T_Q5_K>(segs, dt, via, st);
        case T_Q6_K: return dispatch_dt<T_Q6_K>(segs, dt, via, st);
    }
    return QT_ERR_UNSUPPORTED;
}

}  // namespace gguf
}  // namespace qt

using namespace qt::gguf;

extern "C" {

int qt_gguf_block_elems(int ggml_type) { return block_elems(ggml_type); }
int qt_gguf_block_bytes(int ggml_type) { return block_bytes(ggml_type); }

int qt_gguf_quantize(int ggml_type, const void* src, int src_dtype, int round_via_f16, int64_t nrows,
                     int64_t ncols, void* dst, void* stream) {
    const int be = block_elems(ggml_type);
    if (be < 0) return QT_ERR_UNSUPPORTED;
    if (nrows < 0 || ncols < 0 || ncols % be) return QT_ERR_INVALID;
    if (nrows == 0 || ncols == 0) return QT_OK;
    if (!src || !dst || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return QT_ERR_INVALID;
    const int64_t nblk = nrows * (ncols / be);
    const int tu = tile_units_of(ggml_type);
    const int64_t tiles = (nblk + tu - 1) / tu;
    if (tiles > 2147483647LL) return QT_ERR_INVALID;
    Segs segs{};
    segs.nseg = 1;
    segs.total_tiles = (int)tiles;
    segs.one = Seg{src, (uint8_t*)dst, (long long)nblk};
    return dispatch_type(ggml_type, segs, src_dtype, round_via_f16, (cudaStream_t)stream);
}

// Many tensors of one type in ONE launch.  src[i] / dst[i]: device pointers (16-byte aligned), nelems[i]: element
// count of tensor i (rows are multiples of the block size, so a tensor is a flat array of blocks).
// table_dev: device scratch of at least qt_gguf_batch_table_bytes(n) bytes (the segment table is copied there).
int64_t qt_gguf_batch_table_bytes(int n) { return n <= 0 ? 0 : (int64_t)(((4 * (n + 1) + 15) / 16) * 16) + (int64_t)sizeof(Seg) * n; }

int qt_gguf_quantize_batch(int ggml_type, int n, const void* const* src, const int64_t* nelems, void* const* dst,
                           int src_dtype, int round_via_f16, void* table_dev, void* stream) {
    const int be = block_elems(ggml_type);
    if (be < 0) return QT_ERR_UNSUPPORTED;
    if (n < 0 || (n > 0 && (!src || !nelems || !dst || !table_dev))) return QT_ERR_INVALID;
    if (n == 0) return QT_OK;
    const int tu = tile_units_of(ggml_type);
    const size_t ts_bytes = (size_t)(((4 * (n + 1) + 15) / 16) * 16);
    std::vector<uint8_t> host(ts_bytes + sizeof(Seg) * n);
    int* tile_start = reinterpret_cast<int*>(host.data());
    Seg* table = reinterpret_cast<Seg*>(host.data() + ts_bytes);
    long long tiles = 0;
    int m = 0;                                  // empty tensors are dropped from the table
    for (int i = 0; i < n; i++) {
        if (nelems[i] < 0 || nelems[i] % be) return QT_ERR_INVALID;
        if (nelems[i] == 0) continue;
        if (!src[i] || !dst[i] || ((uintptr_t)src[i] & 15) || ((uintptr_t)dst[i] & 15)) return QT_ERR_INVALID;
        const long long units = nelems[i] / be;
        tile_start[m] = (int)tiles;
        table[m] = Seg{src[i], (uint8_t*)dst[i], units};
        tiles += (units + tu - 1) / tu;
        if (tiles > 2147483647LL) return QT_ERR_INVALID;
        m++;
    }
    if (m == 0) return QT_OK;
    tile_start[m] = (int)tiles;
    cudaStream_t st = (cudaStream_t)stream;
    Segs segs{};
    segs.nseg = m;
    segs.total_tiles = (int)tiles;
    if (m == 1) {
        segs.one = table[0];
    } else {
        // pageable source: the copy is staged before the call returns, so `host` may go out of scope
        cudaError_t e = cudaMemcpyAsync(table_dev, host.data(), host.size(), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { qt::set_last_error("gguf batch table copy", e); return QT_ERR_CUDA; }
        segs.tile_start = reinterpret_cast<const int*>(table_dev);
        segs.table = reinterpret_cast<const Seg*>((const uint8_t*)table_dev + ts_bytes);
    }
    return dispatch_type(ggml_type, segs, src_dtype, round_via_f16, st);
}

int qt_gguf_dequantize(int ggml_type, const void* src, int64_t nrows, int64_t ncols, float* dst, void* stream) {
    const int be = block_elems(ggml_type);
    if (be < 0) return QT_ERR_UNSUPPORTED;
    if (nrows < 0 || ncols < 0 || ncols % be) return QT_ERR_INVALID;
    if (nrows == 0 || ncols == 0) return QT_OK;
    if (!src || !dst || ((uintptr_t)dst & 15)) return QT_ERR_INVALID;
    const int64_t nblk = nrows * (ncols / be);
    cudaStream_t st = (cudaStream_t)stream;
    const uint8_t* s = (const uint8_t*)src;
    if (((uintptr_t)src & 15) == 0) {        // staged kernels: 128-bit loads of the packed bytes
        switch (ggml_type) {
            case T_Q4_0: launch_dequant_staged<T_Q4_0>(s, dst, nblk, st); break;
            case T_Q4_1: launch_dequant_staged<T_Q4_1>(s, dst, nblk, st); break;
            case T_Q5_0: launch_dequant_staged<T_Q5_0>(s, dst, nblk, st); break;
            case T_Q5_1: launch_dequant_staged<T_Q5_1>(s, dst, nblk, st); break;
            case T_Q8_0: launch_dequant_staged<T_Q8_0>(s, dst, nblk, st); break;
            case T_IQ4_NL: launch_dequant_staged<T_IQ4_NL>(s, dst, nblk, st); break;
            case T_Q2_K: launch_dequant_staged<T_Q2_K>(s, dst, nblk, st); break;
            case T_Q3_K: launch_dequant_staged<T_Q3_K>(s, dst, nblk, st); break;
            case T_Q4_K: launch_dequant_staged<T_Q4_K>(s, dst, nblk, st); break;
            case T_Q5_K: launch_dequant_staged<T_Q5_K>(s, dst, nblk, st); break;
            case T_Q6_K: launch_dequant_staged<T_Q6_K>(s, dst, nblk, st); break;
        }
        return qt::check_launch("qt_gguf_dequantize");
    }
    if (be == 32) {
        const int64_t nthreads = nblk * 2;
        const unsigned grid = (unsigned)((nthreads + 255) / 256);
        switch (ggml_type) {
            case T_Q4_0: dequant_simple_kernel<T_Q4_0><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q4_1: dequant_simple_kernel<T_Q4_1><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q5_0: dequant_simple_kernel<T_Q5_0><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q5_1: dequant_simple_kernel<T_Q5_1><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q8_0: dequant_simple_kernel<T_Q8_0><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_IQ4_NL: dequant_simple_kernel<T_IQ4_NL><<<grid, 256, 0, st>>>(s, dst, nblk); break;
        }
    } else {
        const int64_t nthreads = nblk * 8;
        const unsigned grid = (unsigned)((nthreads + 255) / 256);
        switch (ggml_type) {
            case T_Q2_K: dequant_k_kernel<T_Q2_K><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q3_K: dequant_k_kernel<T_Q3_K><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q4_K: dequant_k_kernel<T_Q4_K><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q5_K: dequant_k_kernel<T_Q5_K><<<grid, 256, 0, st>>>(s, dst, nblk); break;
            case T_Q6_K: dequant_k_kernel<T_Q6_K><<<grid, 256, 0, st>>>(s, dst, nblk); break;
        }
    }
    return qt::check_launch("qt_gguf_dequantize");
}

}  // extern "C"
