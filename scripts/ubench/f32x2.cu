// Micro-benchmark: issue rate of scalar vs packed (f32x2) fp32 add/mul on sm_100a, and of FMNMX next to them.
// Answers: does add.rn.f32x2 / mul.rn.f32x2 (no FMA, IEEE per lane) halve the issue slots of a strictly ordered
// fp32 search (K-quant packers)?  Prints warp-instructions per clock per SM and lane-results per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a + i + threadIdx.x;
    unsigned long long* xp = reinterpret_cast<unsigned long long*>(x);
    unsigned long long bb;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {        // 16 scalar FADD
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 1) { // 8 packed add (16 lane results)
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
        } else if (MODE == 2) { // 16 scalar FMUL
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 3) { // 8 packed mul
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
        } else if (MODE == 4) { // 8 FADD + 8 FMUL scalar
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[2 * i]) : "f"(b));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[2 * i + 1]) : "f"(b));
            }
        } else if (MODE == 5) { // 4 packed add + 4 packed mul
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                if (i & 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
                else asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
        } else if (MODE == 6) { // 16 FMNMX
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 7) { // 8 FMNMX + 4 packed add (8 lanes)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                unsigned long long v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(bb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
#pragma unroll
            for (int i = 8; i < 16; i++) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 8) { // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(b));
        } else if (MODE == 9) { // 8 packed FFMA2
#pragma unroll
            for (int i = 0; i < 8; i++) {
                unsigned long long v;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(bb));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warp_instr_per_iter, int lanes_per_iter) {
    int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int ctas = nsm * 4, thr = 256;
    float* out; cudaMalloc(&out, ctas * thr * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<ctas, thr>>>(out, 1.0f, 1.0000001f);
    cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); k<MODE><<<ctas, thr>>>(out, 1.0f, 1.0000001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warps = (double)ctas * thr / 32;
    const double winstr = warps * ITERS * warp_instr_per_iter;
    const double lanes = warps * 32.0 * ITERS * lanes_per_iter;
    printf("%-34s %8.3f ms  %7.2f Gwarp-instr/s/SM  %8.1f Glane-results/s/SM  (%s)\n", name, best, winstr / best / 1e6 / nsm,
           lanes / best / 1e6 / nsm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("16 x add.rn.f32", 16, 16);
    run<1>("8 x add.rn.f32x2", 8, 16);
    run<2>("16 x mul.rn.f32", 16, 16);
    run<3>("8 x mul.rn.f32x2", 8, 16);
    run<4>("8 add + 8 mul scalar", 16, 16);
    run<5>("4 add.f32x2 + 4 mul.f32x2", 8, 16);
    run<6>("16 x min.f32 (FMNMX)", 16, 16);
    run<7>("4 add.f32x2 + 8 FMNMX", 12, 16);
    run<8>("16 x fma.rn.f32", 16, 16);
    run<9>("8 x fma.rn.f32x2", 8, 16);
    return 0;
}
