// Sanitizer driver for the host emulation of the CUDA K-quant / IQ4_NL packers (tests/host_emul.cu, i.e. the
// __host__ __device__ phase functions of quantool_b200/csrc/gguf_kquant.cuh that the kernels run): ordinary, tiny and
// sparse / outlier inputs, output buffers of exactly the packed size.
// Build: nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined
//        -I quantool_b200/csrc -I include host_emul_sanitize.cu host_emul.cu -Xlinker -lasan -Xlinker -lubsan
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

extern "C" {
void emul_iq4_nl(const float*, uint8_t*, int64_t);
void emul_q2_K(const float*, uint8_t*, int64_t);
void emul_q3_K(const float*, uint8_t*, int64_t);
void emul_q4_K(const float*, uint8_t*, int64_t);
void emul_q5_K(const float*, uint8_t*, int64_t);
void emul_q6_K(const float*, uint8_t*, int64_t);
}

int main() {
    const int64_t n = 40 * 1024;
    float* x = (float*)malloc(sizeof(float) * n);
    const struct { void (*pack)(const float*, uint8_t*, int64_t); int elems, bytes; } types[] = {
        {emul_iq4_nl, 32, 18}, {emul_q2_K, 256, 84},  {emul_q3_K, 256, 110},
        {emul_q4_K, 256, 144}, {emul_q5_K, 256, 176}, {emul_q6_K, 256, 210}};
    for (int rep = 0; rep < 3; rep++) {
        for (int64_t i = 0; i < n; i++) {
            const float u = (float)rand() / (float)RAND_MAX - 0.5f;
            x[i] = rep == 0 ? u : rep == 1 ? u * 1e-8f : (i % 97 == 0 ? 1e4f * u : 0.f);
        }
        if (rep == 2)
            for (int i = 0; i < 512; i++) x[i] = 0.f;          // all-zero super-blocks
        for (const auto& t : types) {
            uint8_t* y = (uint8_t*)malloc((size_t)(n / t.elems) * (size_t)t.bytes);
            t.pack(x, y, n / t.elems);
            free(y);
        }
    }
    free(x);
    puts("ok");
    return 0;
}
