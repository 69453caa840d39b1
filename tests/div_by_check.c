/* Host check of the division the lean GPTQ block kernel uses (quantool_b200/csrc/gptq.cu: recip_refined, div_by):
 *     y0 = rcp.approx.ftz(s);  e = fma(-s, y0, 1);  y = fma(y0, e, y0);            once per group
 *     q0 = w * y;  rem = fma(-s, q0, w);  q = fma(y, rem, q0);                      per element
 * The kernel's contract is "the same bits as w / s" (upstream evaluates torch's x / scale).  rcp.approx.ftz.f32 is
 * specified to within 1 ulp of 1/s, so every admissible y0 is one of RN(1/s) - 1 ulp, RN(1/s), RN(1/s) + 1 ulp: all
 * three are tried for each random (w, s) in the range of weights and group scales, with IEEE fmaf, and q is compared
 * with the correctly rounded quotient.  Prints the number of mismatches (expected: 0).
 * Build: gcc -O2 -ffp-contract=off div_by_check.c -lm ; run: ./a.out <pairs> */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float step_ulps(float x, int k) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += (uint32_t)k;
    memcpy(&x, &u, 4);
    return x;
}

static uint64_t state = 88172645463325252ull;
static uint32_t next_u32(void) {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return (uint32_t)(state >> 16);
}

static float random_float(int min_exp, int span) {
    const float m = 1.0f + (float)(next_u32() & 0x7fffff) / 8388608.0f;
    return ldexpf(m, min_exp + (int)(next_u32() % (uint32_t)span));
}

int main(int argc, char** argv) {
    const long pairs = argc > 1 ? atol(argv[1]) : 5000000;
    long mismatches = 0;
    for (long i = 0; i < pairs; i++) {
        const float s = random_float(-20, 24);                 /* scales 1e-6 .. 16 */
        float w = random_float(-24, 30);                       /* |w| 6e-8 .. 64 */
        if (next_u32() & 1) w = -w;
        const float want = w / s;
        const float yr = 1.0f / s;
        for (int k = -1; k <= 1; k++) {
            const float y0 = step_ulps(yr, k);
            const float e = fmaf(-s, y0, 1.0f);
            const float y = fmaf(y0, e, y0);
            const float q0 = w * y;
            const float rem = fmaf(-s, q0, w);
            const float q = fmaf(y, rem, q0);
            mismatches += (q != want);
        }
    }
    printf("%ld\n", mismatches);
    return mismatches != 0;
}
