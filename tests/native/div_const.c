/* div_const.c -- the constant-division identity the hand-written ATMOSPHERE kernel relies on
 * (shaderbox_b200/csrc/native/app_atmosphere_native.h, sbx_div_const): for r = RN(1/d),
 *     q0 = x * r;  e = fma(-d, q0, x);  q = fma(e, r, q0)
 * equals the correctly rounded x / d for every float x with 2^-100 <= |x| <= 2^100.
 * usage: div_const <stride> d...     checks every stride-th fp32 bit pattern (stride 1 = all 2^32, ~1 min on 8 cores)
 * prints one JSON line; exit code 0 = no mismatch inside the range (mismatches outside it are counted, not failures) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], NULL, 10) : 257;
    int k, fail = 0;
    printf("{\"stride\": %llu, \"divisors\": [", (unsigned long long)stride);
    for (k = 2; k < argc; ++k) {
        const float d = (float)atof(argv[k]);
        const float r = 1.0f / d;
        uint64_t i, inside = 0, outside = 0, checked = 0;
        for (i = 0; i < (1ull << 32); i += stride) {
            const uint32_t u = (uint32_t)i;
            float x, want, q0, e, q;
            uint32_t a, b;
            memcpy(&x, &u, 4);
            want = x / d;
            q0 = x * r;
            e = fmaf(-d, q0, x);
            q = fmaf(e, r, q0);
            memcpy(&a, &want, 4);
            memcpy(&b, &q, 4);
            ++checked;
            if (a != b && !(want != want && q != q)) {
                if (fabsf(x) >= 0x1p-100f && fabsf(x) <= 0x1p100f) ++inside; else ++outside;
            }
        }
        printf("%s{\"d\": %g, \"r\": \"%a\", \"checked\": %llu, \"mismatch_in_range\": %llu, \"mismatch_outside\": %llu}", k > 2 ? ", " : "",
               d, r, (unsigned long long)checked, (unsigned long long)inside, (unsigned long long)outside);
        if (inside) fail = 1;
    }
    printf("], \"ok\": %s}\n", fail ? "false" : "true");
    return fail;
}
