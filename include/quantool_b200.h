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
#define QT_GGML_Q4_K 12
#define QT_GGML_Q5_K 13
#define QT_GGML_Q6_K 14

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
int qt_gguf_dequantize(int ggml_type, const void* src, int64_t nrows, int64_t ncols, float* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUANTOOL_B200_H */
