/* Checks the identity k_tile's exact_div relies on (draw_b200/csrc/device_math.cuh):
 *     q0 = RN(e * rf),  r = fma(-f, q0, e),  q = fma(r, rf, q0)   with rf = RN(1 / f)
 * equals the IEEE quotient RN(e / f) (Markstein).  Inputs: (1) random integers and random mantissas
 * in the range TRI_FASTDIV admits (1 <= f <= 2^40), (2) the hardest cases for rounding: quotients
 * within ~2^-49 (relative) of a midpoint between two floats, built by modular inversion of f's
 * significand.  Test infrastructure only; prints "tested=<n> bad=<m>".
 * build: gcc -O2 -mfma -ffp-contract=off fastdiv_check.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline uint64_t rng(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline float exact_div(float e, float f, float rf) {
    const float q0 = e * rf;
    const float r = fmaf(-f, q0, e);
    return fmaf(r, rf, q0);
}
static long bad = 0, tested = 0;
static void check(float e, float f) {
    const float want = e / f, got = exact_div(e, f, 1.0f / f);
    tested++;
    if (memcmp(&got, &want, 4)) {
        if (bad < 10) printf("mismatch e=%a f=%a got %a want %a\n", e, f, got, want);
        bad++;
    }
}
int main(int argc, char **argv) {
    uint64_t s = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    const long n = argc > 2 ? atol(argv[2]) : 1000000;
    for (long i = 0; i < n; i++) {
        const uint64_t a = rng(&s), b = rng(&s);
        /* (1a) integers, as edge functions of snapped vertices are */
        const int k = 1 + (int)((a >> 3) % 34), m = 1 + (int)((a >> 9) % 40);
        check((float)(double)((b >> 8) % (1ull << k)), (float)(double)(1 + ((b >> 20) % (1ull << m))));
        /* (1b) random significands */
        uint32_t ue = ((127u + (uint32_t)((a >> 20) % 60)) << 23) | (uint32_t)(b & 0x7FFFFF);
        uint32_t uf = ((127u + (uint32_t)((a >> 30) % 40)) << 23) | (uint32_t)((b >> 23) & 0x7FFFFF);
        float e, f;
        memcpy(&e, &ue, 4);
        memcpy(&f, &uf, 4);
        check(e, f);
        /* (2) near-midpoint quotients: (2Q+1) * mf = d (mod 2^25) for small odd d */
        const uint64_t mf = (rng(&s) & 0x7FFFFF) | 0x800000 | 1;
        uint64_t inv = mf;
        for (int it = 0; it < 6; it++) inv = (inv * (2 - mf * inv)) & 0x1FFFFFF;
        for (int d = -15; d <= 15; d += 2) {
            const uint64_t t = ((uint64_t)((int64_t)d * (int64_t)inv)) & 0x1FFFFFF;
            if (t < (1u << 24)) continue;
            const __int128 num = (__int128)t * mf - d;
            const uint64_t me = (uint64_t)(num >> 25);
            if ((num & 0x1FFFFFF) || me == 0 || me >= (1u << 24)) continue;
            check((float)me, (float)mf);
            check(ldexpf((float)me, 9), ldexpf((float)mf, 13));
        }
    }
    printf("tested=%ld bad=%ld\n", tested, bad);
    return bad != 0;
}
