/*
 * quantool_b200 C-ABI — the drop-in boundary underneath the quantool method plugins.
 *
 * The reference (langtech-bsc/quantool) has no FFI of its own: its plugins hand the work to
 * llm-compressor (in-process PyTorch) and llama.cpp (child process).  Each entry point below
 * names the reference call site whose arithmetic it replaces (ref/ = /root/reference/,
 * UPSTREAM = un-vendored dependency restated in SURVEY.md §A-§D, CT = compressed_tensors).
 *
 * Conventions (last row):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless marked host
 *   - the caller owns all buffers; nothing is allocated or freed behind the ABI
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no hidden syncs
 *   - return 0 on success, <0 on error (QT_ERR_*); never throws; qt_last_error() gives text
 *   - thread-safe for distinct streams
 */
#ifndef QUANTOOL_B200_H
#define QUANTOOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QT_OK 0
#define QT_ERR_INVALID (-1)
#define QT_ERR_CUDA (-2)
#define QT_ERR_UNSUPPORTED (-3)

/* element types of caller buffers */
#define QT_F32 0
#define QT_F16 1
#define QT_BF16 2

/* ggml tensor types (gguf-py GGMLQuantizationType values) */
#define QT_GGML_Q4_0 2
#define QT_GGML_Q4_1 3
#define QT_GGML_Q5_0 6
#define QT_GGML_Q5_1 7
#define QT_GGML_Q8_0 8
#define QT_GGML_Q2_K 10
#define QT_GGML_Q3_K 11
#define QT_GGML_Q4_K 12
#define QT_GGML_Q5_K 13
#define QT_GGML_Q6_K 14
#define QT_GGML_IQ4_NL 20

const char* qt_last_error(void);
int qt_abi_version(void);
int qt_device_sm_count(void);

/* ---- GGUF block packing ------------------------------------------------------------------
 * Replaces the `llama-quantize` child process: ref/src/quantool/methods/llama_cpp/llama_cpp.py:165-178
 * (GGUF._quantize_gguf -> run_command), i.e. UPSTREAM llama.cpp ggml-quants.c
 * quantize_row_<type>_ref / dequantize_row_<type>. */
int qt_gguf_block_elems(int ggml_type);   /* 32 or 256; -1 if unsupported */
int qt_gguf_block_bytes(int ggml_type);
/* src: [nrows, ncols] row-major of src_dtype; dst: nrows * ncols/block_elems * block_bytes bytes.
 * round_via_f16 != 0 rounds each element through fp16 first, as the reference's
 * HF -> model.f16.gguf -> llama-quantize chain does (llama_cpp.py:207,238). */
int qt_gguf_quantize(int ggml_type, const void* src, int src_dtype, int round_via_f16, int64_t nrows,
                     int64_t ncols, void* dst, void* stream);
/* Many tensors of one type in ONE launch (a model is hundreds of small tensors; llama-quantize walks them in a
 * loop, llama_cpp.py:165-178).  src[i]/dst[i]: device pointers, nelems[i]: elements of tensor i; table_dev: device
 * scratch of qt_gguf_batch_table_bytes(n) bytes. */
int64_t qt_gguf_batch_table_bytes(int n);
int qt_gguf_quantize_batch(int ggml_type, int n, const void* const* src, const int64_t* nelems, void* const* dst,
                           int src_dtype, int round_via_f16, void* table_dev, void* stream);
int qt_gguf_dequantize(int ggml_type, const void* src, int64_t nrows, int64_t ncols, float* dst, void* stream);

unsigned long long qt_launch_count(void);   /* kernel launches issued through this library since load */

/* ---- compressed-tensors primitives -----------------------------------------------------------
 * What the reference's artifacts are made of when llm-compressor plugins save with
 * save_compressed=True (ref/src/quantool/methods/llm_compressor/base.py:188,230):
 *   CT/quantization/utils/helpers.py:50-137      calculate_qparams (MinMax observer)
 *   CT/quantization/lifecycle/forward.py:36-73   quantize -> int8 codes / dequantize
 *   CT/compressors/pack_quantized/helpers.py:20-161  pack_to_int32 / unpack_from_int32
 * Arithmetic is evaluated in `dtype` (bf16/fp16 tensors round after every op, as torch does). */
/* W [N,K] of `dtype`; group_size 0 = one scale per row; scale, zp: fp32 [N, K/group_size] (zp integer-valued) */
int qt_minmax_qparams(const void* W, int dtype, int N, int K, int group_size, int num_bits, int symmetric,
                      float* scale, float* zp, void* stream);
/* scale [N,G] of `dtype`; zp fp32 [N,G] or NULL; g_idx int32 [K] or NULL (column -> group);
 * codes int8 [N,K] and/or dq_out [N,K] of `dtype` (either may be NULL, not both) */
int qt_quantize_codes(const void* W, const void* scale, const float* zp, const int* g_idx, int dtype, int N, int K,
                      int G, int group_size, int num_bits, int8_t* codes, void* dq_out, void* stream);
int qt_pack_int32(const int8_t* codes, int N, int K, int num_bits, int32_t* packed, void* stream);
int qt_unpack_int32(const int32_t* packed, int N, int K, int num_bits, int8_t* codes, void* stream);

/* ---- GPTQ ---------------------------------------------------------------------------------------
 * Replaces UPSTREAM llmcompressor modifiers/quantization/gptq/gptq_quantize.py (accumulate_hessian,
 * quantize_weight), which the reference reaches by building GPTQModifier at
 * ref/src/quantool/methods/llm_compressor/gptq/gptq.py:86 and running llmcompressor.oneshot at
 * ref/src/quantool/methods/llm_compressor/base.py:159-161. */
/* H fp32 [K,K] (raw sums, caller zeroes it first) += X^T X on upper-triangle tiles; X [T,K] of `dtype` = QT_BF16 or
 * QT_F16 (the model's activation dtype goes to the tensor cores unchanged; QT_F32 -> QT_ERR_UNSUPPORTED); tcgen05 */
int qt_hessian_accumulate(const void* X, int dtype, int64_t T, int K, float* H, void* stream);
/* H <- factor * H (factor = 2 / n_samples) on the upper triangle, mirrored to the lower */
int qt_hessian_finalize(float* H, int K, float factor, void* stream);
/* Exact diagonal: the tensor-core accumulator truncates, which biases the sums of squares on the diagonal by
 * ~-5e-6 and can reorder argsort(diag H) (the act_order permutation).  diag[c] += sum_t X[t][c]^2 in fp32
 * round-to-nearest, deterministic; scratch = 32*K floats.  qt_hessian_set_diagonal writes the raw sums into H
 * (before finalize / before a cross-rank reduction). */
int qt_hessian_diag_accumulate(const void* X, int dtype, int64_t T, int K, float* diag, float* scratch, void* stream);
int qt_hessian_set_diagonal(float* H, int K, const float* diag, void* stream);
int qt_hessian_set_splits(int splits);   /* tuning: force the token split count (0 = heuristic) */
/* Leave n SMs to other streams for the following qt_hessian_accumulate calls (0 = take every SM): used while the
 * NCCL all-reduce of the previous input's Hessian is in flight - its CTAs cannot be placed next to the SYRK's. */
int qt_hessian_reserve_sms(int n);
/* Packed upper block-triangle of a K x K fp32 matrix: what the multi-GPU path puts on NVLink instead of the full
 * square ("all-reduce the packed triangle").  Row block i (128 rows) is stored from column
 * floor(128 i / align) * align; align = 256 for the raw Hessian sums (the SYRK's tile width), 128 for the
 * upper-triangular factor U.  zero_below != 0 also clears everything left of the stored part (U on receivers). */
int64_t qt_tri_packed_elems(int K, int align);
int qt_tri_pack(const float* M, int K, int align, float* packed, void* stream);
int qt_tri_unpack(const float* packed, int K, int align, int zero_below, float* M, void* stream);
/* dead[i] = (H[i][i]==0); damp = percdamp*mean(diag); Hf = J P^T (H' + damp I) P J (lower triangle),
 * perm int32 [K] or NULL, dead uint8 [K], damp_scratch fp32 [1] */
int qt_gptq_prepare_hessian(const float* H, const int* perm, int K, float percdamp, float* Hf, uint8_t* dead,
                            float* damp_scratch, void* stream);
/* in place: A = Hf -> U = cholesky(H^-1, upper); X, W: [K,K] fp32 scratch; info: device int32,
 * 0 = ok else 1-based failing pivot (upstream's LinAlgError path: caller sets U = I) */
int qt_gptq_hinv_factor(float* A, float* X, float* W, int K, int* info, void* stream);
/* Same contract with the big products on the tensor cores (3xTF32, fp32-faithful): four more K x K fp32
 * workspaces (tf32 splits of L and of the inverse blocks); K % 256 == 0.  stages: bit 0 = Cholesky trailing updates,
 * bit 1 = triangular-inverse merges on the tensor cores (3 = both). */
int qt_gptq_hinv_factor_tc(float* A, float* X, float* W, float* Lh, float* Ll, float* Dh, float* Dl, int K, int* info,
                           int stages, void* stream);
int qt_set_identity(float* U, int K, void* stream);
int qt_sgemm(int b_is_nk, const float* A, const float* B, float* C, int M, int N, int Kd, int lda, int ldb, int ldc,
             float alpha, float beta, int lower_tiles_only, int a_lower_tri, int b_lower_tri, void* stream);
int qt_sgemm_ex(int b_is_nk, const float* A, const float* B, float* C, int M, int N, int Kd, int lda, int ldb, int ldc,
                float alpha, float beta, int lower_tiles_only, int a_lower_tri, int b_lower_tri, int tri_row_offset,
                void* stream);
/* pieces of the multi-GPU inverse-factor chain: Cholesky + triangular inverse of an n x n block inside
 * buffers of leading dimension ld (info accumulates), and U[i][j] = X[K-1-i][K-1-j] (upper) */
int qt_tri_chain_block(float* A, float* X, float* W, int n, int ld, int* info, void* stream);
int qt_flip_upper(const float* X, float* U, int K, void* stream);
/* Wp[n][j] = float(W[n][perm[j]]) with dead columns zeroed; out[n][c] = cast(Wp[n][inv_perm[c]]) */
int qt_gptq_permute_in(const void* W, int dtype, const int* perm, const uint8_t* dead, float* Wp, int N, int K,
                       void* stream);
int qt_gptq_permute_out(const float* Wp, const int* inv_perm, void* out, int dtype, int N, int K, void* stream);
/* tuning / A-B checks: 1 = run every 128-column block through the generic block kernel (also used for partial
 * blocks), 0 (default) = the lean full-block kernel.  Both produce the same bits. */
int qt_gptq_set_block_kernel(int generic);
/* blocked column loop (block 128).  mode 0: re-fit group qparams at group starts (group_size 32/64/128),
 * 1: static scales looked up through g_idx (actorder=weight), 2: one scale per row.  W in/out.
 * U_hi/U_lo (qt_split_tf32_transpose of U) non-NULL: lazy update on the tensor cores, batched over 512-column
 * outer blocks, err_scratch [2,N,512]; NULL: fp32 FFMA GEMM per block, err_scratch [N,128]. */
int qt_gptq_quantize_weight(float* W, const float* U, const float* U_hi, const float* U_lo, float* err_scratch,
                            float* scale, float* zp, const int* g_idx, float* losses, int N, int K, int G,
                            int group_size, int num_bits, int symmetric, int mode, void* stream);
/* tf32 operand splits of the 3xTF32 tensor-core GEMMs (qt_gemm_tf32x3): x -> hi (tf32-exact) + lo */
int qt_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* ut_hi + ut_lo = U^T (the K-major B operand of the lazy-batch update W[:, i2:] -= Err * U[i1:i2, i2:]) */
int qt_split_tf32_transpose(const float* U, float* ut_hi, float* ut_lo, int K, void* stream);

/* ---- SmoothQuant / AWQ --------------------------------------------------------------------------
 * Replaces UPSTREAM llmcompressor SmoothQuantModifier built at
 * ref/src/quantool/methods/llm_compressor/smoothquant/smoothquant.py:77-84 and AWQModifier
 *  built at ref/src/quantool/methods/llm_compressor/awq/awq.py:81. */
int qt_fill_f32(float* p, int n, float v, void* stream);
/* running per-channel min/max over rows of X [T,K]; mn/mx fp32 [K] initialised to +/-FLT_MAX */
int qt_channel_minmax(const void* X, int dtype, int64_t T, int K, float* mn, float* mx, void* stream);
/* sum[c] += sum_t |X[t][c]| */
int qt_channel_abs_sum(const void* X, int dtype, int64_t T, int K, float* sum, void* stream);
/* s = (amax-amin)^alpha / (2 max|w|)^(1-alpha), where(w>0, s, act), max(s, 1e-5); evaluated in compute_dtype */
int qt_smooth_scales(const float* amin, const float* amax, const float* wmin, const float* wmax, float alpha,
                     float one_minus_alpha, int compute_dtype, float* s_out, int K, void* stream);
/* in place W[n][c] *= s[c] (divide: /=), or by_row: s[n] */
int qt_scale_matrix(void* W, int dtype, int N, int K, const float* s, int divide, int by_row, void* stream);
/* colsum[c] += sum_n |W[n][c]| / (group amax + 1e-6); gamax_scratch fp32 [N, K/group_size] */
int qt_awq_wmean_accumulate(const void* W, int dtype, int N, int K, int group_size, float* gamax_scratch,
                            float* colsum, void* stream);
/* out = pseudo_quantize(W * s) / s in one pass (AWQ grid-search candidate) */
int qt_awq_scale_qdq(const void* W, int dtype, int N, int K, const float* s, int group_size, int num_bits,
                     int symmetric, void* out, void* stream);
/* Gram-form reconstruction loss for a single-Linear AWQ parent (o_proj / down_proj mappings; UPSTREAM AWQModifier
 * _compute_loss on F.linear outputs, SURVEY.md B.3 and section 7 hard part 6):
 *   || X D^T ||_F^2 = tr(D G D^T),  G = X^T X (qt_hessian_accumulate + qt_hessian_finalize(.., 1.0), cast to bf16),
 *   D = pseudo_quant(W * s) / s - W.
 * qt_awq_scale_qdq_delta writes D for one grid point (bf16 operand + fp32 copy, group_size 32 / 64 / 128);
 * qt_awq_gram_loss runs ONE kind::f16 tcgen05 GEMM (D * G, fp32 in TMEM) whose epilogue multiplies by the fp32 D
 * tile and reduces into *loss (device double, += ): neither the candidate output nor D*G is materialised. */
int qt_awq_scale_qdq_delta(const void* W, int dtype, int N, int K, const float* s, int group_size, int num_bits,
                           int symmetric, void* delta_bf16, float* delta_f32, void* stream);
int qt_awq_gram_loss(const void* D_bf16, const float* D_f32, const void* G_bf16, int M, int K, double* loss, void* stream);

/* *out (device double) += sum (a-b)^2 */
int qt_sq_err_sum(const void* a, const void* b, int dtype, int64_t n, double* out, void* stream);

/* fp32-faithful tensor-core GEMM of the inverse-Hessian chain (3xTF32, tcgen05): C (op)= +-A B^T with
 * A = a_hi + a_lo [M,Kd], B = b_hi + b_lo [N,Kd] (qt_split_tf32 outputs), row-major, contraction over columns.
 * flags: bit0 negate, bit1 accumulate into C (TMA reduce-add), bit2 lower tiles only, bits 4-5 / 6-7: A / B
 * triangular (1 lower, 2 upper; the k range is trimmed).  Kd % 32 == 0.  Stands in for the cuBLAS fp32 GEMMs
 * inside torch.linalg.cholesky / cholesky_inverse of UPSTREAM gptq_quantize.py. */
int qt_gemm_tf32x3(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, float* C, int M, int N,
                   int Kd, int lda, int ldb, int ldc, int flags, void* stream);

/* ---- calibration forward, elementwise pieces ----------------------------------------------------
 * The reference runs calibration data through the HF model inside llm-compressor `oneshot`
 * (ref/src/quantool/methods/llm_compressor/base.py:162); these are one-pass replacements for the
 * normalisation / rotary / gated-activation modules between the GEMMs of that forward
 * (transformers LlamaRMSNorm, apply_rotary_pos_emb, LlamaMLP), with the same rounding points. */
/* out[t] = weight * (x[t].float() * rsqrt(mean(x[t]^2) + eps)).to(dtype); x, out: [T, H], H % 8 == 0 */
int qt_rms_norm(const void* x, const void* weight, void* out, int dtype, int64_t T, int H, float eps, void* stream);
/* in place x = x*cos + rotate_half(x)*sin on x[T, n_heads*head_dim]; cos/sin [seq, head_dim]; position = t % seq */
int qt_rope_inplace(void* x, const void* cos_t, const void* sin_t, int dtype, int64_t T, int seq, int n_heads,
                    int head_dim, void* stream);
/* out = silu(gate) * up, n elements, n % 8 == 0 */
int qt_silu_mul(const void* gate, const void* up, void* out, int dtype, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUANTOOL_B200_H */
