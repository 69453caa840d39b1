/* Sanitizer driver for the C oracle (oracle/ggml_quants.c): every block type on ordinary, tiny, and sparse /
 * outlier inputs, output buffers of exactly the packed size, 1 / 2 / all threads.
 * Build: gcc -O1 -g -fsanitize=address,undefined -ffp-contract=off -pthread oracle_sanitize.c ../oracle/ggml_quants.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

int oracle_block_elems(int t);
int oracle_block_bytes(int t);
int oracle_quantize(int t, const float* x, void* y, int64_t nrows, int64_t ncols);
int oracle_dequantize(int t, const void* x, float* y, int64_t nrows, int64_t ncols);
void oracle_set_threads(int n);

int main(void) {
    const int types[] = {2, 3, 6, 7, 8, 10, 11, 12, 13, 14, 20}; /* Q4_0 Q4_1 Q5_0 Q5_1 Q8_0 Q2_K..Q6_K IQ4_NL */
    const int64_t nrows = 7, ncols = 1024;
    float* x = malloc(sizeof(float) * nrows * ncols);
    float* z = malloc(sizeof(float) * nrows * ncols);
    for (int rep = 0; rep < 3; rep++) {
        for (int64_t i = 0; i < nrows * ncols; i++) {
            const float u = (float)rand() / (float)RAND_MAX - 0.5f;
            x[i] = rep == 0 ? u : rep == 1 ? u * 1e-8f : (i % 97 == 0 ? 1e4f * u : 0.0f);
        }
        if (rep == 2)
            for (int64_t i = 0; i < 256; i++) x[i] = 0.0f; /* an all-zero super-block */
        oracle_set_threads(rep); /* 0 = all cores */
        for (unsigned k = 0; k < sizeof(types) / sizeof(types[0]); k++) {
            const int t = types[k], be = oracle_block_elems(t), bb = oracle_block_bytes(t);
            void* y = malloc((size_t)nrows * (size_t)(ncols / be) * (size_t)bb);
            if (oracle_quantize(t, x, y, nrows, ncols) || oracle_dequantize(t, y, z, nrows, ncols)) return 2;
            for (int64_t i = 0; i < nrows * ncols; i++)
                if (!isfinite(z[i])) return 3;
            free(y);
        }
    }
    free(x);
    free(z);
    puts("ok");
    return 0;
}
