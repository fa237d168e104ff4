/* math_vs_libm.c -- host check of shaderbox_b200/include/sbx/sbx_math.h against the platform libm.
 * usage: math_vs_libm <fn> <stride> [nthreads]    fn in sinf cosf expf powf tanf acosf atan2f sincosf
 * Walks every `stride`-th fp32 bit pattern (stride 1 = exhaustive) and prints the number of
 * results whose bits differ from libm's (NaNs compare equal to NaNs).                         */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../shaderbox_b200/include/sbx/sbx_math.h"

typedef struct { int fn; unsigned stride; unsigned tid, nth; unsigned long long bad, n; unsigned first_bad; } job_t;

static unsigned long long rng_state = 88172645463325252ull;
static int same(float a, float b) { return (isnan(a) && isnan(b)) || sbx_f2u(a) == sbx_f2u(b); }

static void* run(void* arg) {
    job_t* j = (job_t*)arg;
    j->bad = 0; j->n = 0; j->first_bad = 0;
    unsigned long long total = 0x100000000ull / j->stride;
    for (unsigned long long k = j->tid; k < total; k += j->nth) {
        unsigned u = (unsigned)(k * j->stride);
        float x = sbx_u2f(u), a, b;
        switch (j->fn) {
            case 0: a = sbx_sinf(x); b = sinf(x); break;
            case 1: a = sbx_cosf(x); b = cosf(x); break;
            case 2: a = sbx_expf(x); b = expf(x); break;
            case 4: a = sbx_tanf(x); b = tanf(x); break;
            case 5: a = sbx_acosf(x); b = acosf(x); break;
            case 6: a = sbx_atanf(x); b = atanf(x); break;
            case 8: {   /* the fused pair against libm's sinf and cosf: report a sine mismatch, else the cosine */
                float c;
                sbx_sincosf(x, &a, &c);
                b = sinf(x);
                if (same(a, b)) { a = c; b = cosf(x); }
                break;
            }
            case 7: {
                unsigned long long h = (k + 1) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
                float y;
                switch (h & 3) {
                    case 0: y = sbx_u2f((unsigned)(h >> 32)); break;
                    case 1: y = ((float)((h >> 35) & 0xffff) / 4096.0f) - 8.0f; break;
                    case 2: y = x * (((float)((h >> 35) & 0xffff) / 16384.0f) - 2.0f); break;
                    default: y = 1.0f; break;
                }
                if (h & 4) { a = sbx_atan2f(x, y); b = atan2f(x, y); } else { a = sbx_atan2f(y, x); b = atan2f(y, x); }
                break;
            }
            default: {
                /* powf: x from the walk, y from a hash of k over a mix of interesting ranges */
                unsigned long long h = (k + 1) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
                float y;
                switch (h & 7) {
                    case 0: y = 1.0f / 2.2f; break;
                    case 1: y = 1.5f; break;
                    case 2: y = 10.0f; break;
                    case 3: y = 1500.0f; break;
                    case 4: y = 30.0f; break;
                    case 5: y = sbx_u2f((unsigned)(h >> 32)); break;
                    case 6: y = (float)((int)((h >> 40) & 63) - 32); break;
                    default: y = ((float)((h >> 35) & 0xffff) / 4096.0f) - 8.0f; break;
                }
                a = sbx_powf(x, y); b = powf(x, y);
            }
        }
        j->n++;
        if (!same(a, b)) { if (!j->bad) j->first_bad = u; j->bad++; }
    }
    return 0;
}

int main(int argc, char** argv) {
    (void)rng_state;
    if (argc < 3) { fprintf(stderr, "usage: %s fn stride [nthreads]\n", argv[0]); return 2; }
    const char* names[] = {"sinf", "cosf", "expf", "powf", "tanf", "acosf", "atanf", "atan2f", "sincosf"};
    int fn = -1;
    for (int i = 0; i < 9; ++i) if (!strcmp(argv[1], names[i])) fn = i;
    if (fn < 0) { fprintf(stderr, "unknown fn\n"); return 2; }
    unsigned stride = (unsigned)strtoul(argv[2], 0, 10);
    unsigned nth = argc > 3 ? (unsigned)atoi(argv[3]) : 8;
    pthread_t th[64]; job_t jobs[64];
    for (unsigned t = 0; t < nth; ++t) { jobs[t] = (job_t){fn, stride, t, nth, 0, 0, 0}; pthread_create(&th[t], 0, run, &jobs[t]); }
    unsigned long long bad = 0, n = 0; unsigned first = 0;
    for (unsigned t = 0; t < nth; ++t) { pthread_join(th[t], 0); bad += jobs[t].bad; n += jobs[t].n; if (jobs[t].bad && !first) first = jobs[t].first_bad; }
    printf("{\"fn\": \"%s\", \"checked\": %llu, \"mismatch\": %llu, \"first_bad_bits\": \"0x%08x\"}\n", names[fn], n, bad, first);
    return bad ? 1 : 0;
}
