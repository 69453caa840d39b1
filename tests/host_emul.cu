// Host emulation of the K-quant CUDA kernels: runs the SAME __host__ __device__ phase
// functions (quantool_b200/csrc/gguf_kquant.cuh) in a loop over thread ids, with the shared
// struct on the heap and a full "barrier" between phases.  Lets the not-gpu test suite check
// the device arithmetic against the oracle on the CPU.  Test infrastructure, not product.
#include <stdlib.h>
#include <vector>

#include "gguf_kquant.cuh"

using namespace qt::kq;

template <int BB, bool Q5>
static void emul_k45(const float* x, uint8_t* out, int64_t nsuper) {
    constexpr int NSB = 32;
    using S = K45Shared<NSB, BB>;
    S* s = (S*)malloc(sizeof(S));
    std::vector<K45Thread> th(NSB * 8);
    for (int64_t base = 0; base < nsuper; base += NSB) {
        const int nvalid = (int)((nsuper - base) < NSB ? (nsuper - base) : NSB);
        for (int e = 0; e < NSB * 256; e++)
            s->u.x[e / 32][e % 32] = (e < nvalid * 256) ? x[base * 256 + e] : 0.f;
        for (int t = 0; t < NSB * 8; t++) {
            if (Q5) k45_phase_a(t, *s, th[t], 31, -0.5f, 0.1f, 15);
            else    k45_phase_a(t, *s, th[t], 15, -1.f, 0.1f, 20);
        }
        for (int t = 0; t < NSB * 8; t++) k45_phase_b<S, BB>(t, *s, th[t], Q5 ? 31 : 15);
        for (int t = 0; t < NSB * 8; t++) { if (Q5) q5k_phase_c(t, *s); else q4k_phase_c(t, *s); }
        memcpy(out + base * BB, s->u.o.out, (size_t)nvalid * BB);
    }
    free(s);
}

template <int BB, bool Q3>
static void emul_k23(const float* x, uint8_t* out, int64_t nsuper) {
    constexpr int NSB = 16;
    using S = K23Shared<NSB, BB>;
    S* s = (S*)malloc(sizeof(S));
    std::vector<K2Thread> t2(NSB * 16);
    std::vector<K3Thread> t3(NSB * 16);
    for (int64_t base = 0; base < nsuper; base += NSB) {
        const int nvalid = (int)((nsuper - base) < NSB ? (nsuper - base) : NSB);
        for (int e = 0; e < NSB * 256; e++)
            s->x[e / 16][e % 16] = (e < nvalid * 256) ? x[base * 256 + e] : 0.f;
        for (int t = 0; t < NSB * 16; t++) { if (Q3) q3k_phase_a(t, *s, t3[t]); else q2k_phase_a(t, *s, t2[t]); }
        for (int t = 0; t < NSB * 16; t++) { if (Q3) q3k_phase_b(t, *s, t3[t]); else q2k_phase_b(t, *s, t2[t]); }
        for (int t = 0; t < NSB * 16; t++) { if (Q3) q3k_phase_c(t, *s); else q2k_phase_c(t, *s); }
        memcpy(out + base * BB, s->out, (size_t)nvalid * BB);
    }
    free(s);
}

extern "C" {
void emul_iq4_nl(const float* x, uint8_t* out, int64_t nblocks) {
    static const float tbl[16] = QT_IQ4NL_VALUES;
    for (int64_t b = 0; b < nblocks; b++) {
        float v[32];
        for (int j = 0; j < 32; j++) v[j] = x[b * 32 + j];
        iq4nl_block(tbl, v, out + b * 18);
    }
}
void emul_q2_K(const float* x, uint8_t* out, int64_t nsuper) { emul_k23<84, false>(x, out, nsuper); }
void emul_q3_K(const float* x, uint8_t* out, int64_t nsuper) { emul_k23<110, true>(x, out, nsuper); }
void emul_q4_K(const float* x, uint8_t* out, int64_t nsuper) { emul_k45<144, false>(x, out, nsuper); }
void emul_q5_K(const float* x, uint8_t* out, int64_t nsuper) { emul_k45<176, true>(x, out, nsuper); }
void emul_q6_K(const float* x, uint8_t* out, int64_t nsuper) {
    constexpr int NSB = 16;
    using S = K6Shared<NSB>;
    S* s = (S*)malloc(sizeof(S));
    std::vector<K6Thread> th(NSB * 16);
    for (int64_t base = 0; base < nsuper; base += NSB) {
        const int nvalid = (int)((nsuper - base) < NSB ? (nsuper - base) : NSB);
        for (int e = 0; e < NSB * 256; e++)
            s->x[e / 16][e % 16] = (e < nvalid * 256) ? x[base * 256 + e] : 0.f;
        for (int t = 0; t < NSB * 16; t++) q6k_phase_a(t, *s, th[t]);
        for (int t = 0; t < NSB * 16; t++) q6k_phase_b(t, *s, th[t]);
        for (int t = 0; t < NSB * 16; t++) q6k_phase_c(t, *s);
        memcpy(out + base * 210, s->out, (size_t)nvalid * 210);
    }
    free(s);
}
}
