// Exhaustive CPU check of the reciprocal-multiply + FMA-correction quotient used by the CUDA kernels
// against the reference's "fp32 true division, round to 16-bit" (quant_funcs.py:37 `tensor / scales`).
// For every finite w >= 0 and every finite s > 0 of the 16-bit dtype:
//     want = rD( w / s )                       (IEEE fp32 division, then RNE to D)
//     r = 1/s (fp32, IEEE);  t = w*r;  e = fma(-s, t, w);  q = fma(r, e, t);  got = rD(q)
// Build: gcc -O2 -mf16c -mfma -fopenmp div_check.c -o div_check -lm ; usage: ./div_check f16|bf16
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <immintrin.h>

static inline float h2f(uint16_t h) { return _cvtsh_ss(h); }
static inline uint16_t f2h(float f) { return _cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC); }
static inline float b2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t f2b(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

int main(int argc, char** argv) {
    int bf = argc > 1 && !strcmp(argv[1], "bf16");
    uint32_t top = bf ? 0x7f80u : 0x7c00u;
    long long bad = 0, total = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(dynamic, 64)
    for (uint32_t sb = 1; sb < top; ++sb) {
        float s = bf ? b2f((uint16_t)sb) : h2f((uint16_t)sb);
        float r = 1.0f / s;
        if (bf && (!isfinite(r) || r == 0.0f || s < 1e-30f || s > 1e30f)) continue;  // scales live in [~4e-8, 3e38/..]; keep the sane range
        for (uint32_t wb = 0; wb < top; ++wb) {
            float w = bf ? b2f((uint16_t)wb) : h2f((uint16_t)wb);
            float want32 = w / s;
            float t = w * r;
            float e = fmaf(-s, t, w);
            float q = fmaf(r, e, t);
            uint16_t a = bf ? f2b(want32) : f2h(want32);
            uint16_t b = bf ? f2b(q) : f2h(q);
            if (bf && (!isfinite(want32) || fabsf(want32) < 1e-30f)) continue;
            total++;
            if (a != b) {
                bad++;
                if (bad < 10) printf("mismatch w=%04x s=%04x want=%04x got=%04x (%.9g / %.9g)\n", wb, sb, a, b, w, s);
            }
        }
    }
    printf("%s pairs=%lld mismatches=%lld\n", bf ? "bf16" : "f16", total, bad);
    return bad != 0;
}
