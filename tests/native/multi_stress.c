/* multi_stress.c -- drives the single-process N-GPU group (sbx_multi_*) through thousands of small frames.  Built by
 * tests/test_host_mock_cpu.py together with libsbx's sources and the mock driver (tests/native/fake_cuda.c) under
 * ThreadSanitizer: the hand-off between the calling thread and the per-GPU worker threads (task slot, go / finished
 * counters, status array, condition variable) must be free of data races, lost wake-ups and deadlocks.  Nothing is
 * rendered here -- the mock records launches -- so this says nothing about pixels; tests/native/multi_c.c does that
 * on real GPUs.
 * usage: multi_stress <n_gpus> <frames>      prints one JSON line; exit code 0 = every call returned SBX_OK */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../../include/sbx.h"

#define FAIL(...) do { printf("{\"ok\": false, \"error\": \""); printf(__VA_ARGS__); printf("\"}\n"); return 1; } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 8, frames = argc > 2 ? atoi(argv[2]) : 2000;
    sbx_multi* group = NULL;
    sbx_params p;
    void* pinned = NULL;
    float* pageable;
    float* dev = NULL;
    float ms[64];
    int f, st;
    struct timespec nap = {0, 2000000};   /* 2 ms: longer than the workers spin, so they go to sleep on the condition variable */

    if (sbx_device_count() < n) FAIL("needs %d (mock) GPUs", n);
    if ((st = sbx_multi_create(NULL, n, &group)) != SBX_OK) FAIL("sbx_multi_create: %s", sbx_strerror(st));
    if ((st = sbx_multi_load_app(group, "APP_CLOUDS", NULL)) != SBX_OK) FAIL("sbx_multi_load_app: %s", sbx_multi_last_error(group));
    if (sbx_default_params(&p, 96, 64) != SBX_OK) FAIL("sbx_default_params");
    if ((st = sbx_host_alloc(sbx_multi_ctx(group, 0), (size_t)96 * 64 * 16, &pinned)) != SBX_OK) FAIL("sbx_host_alloc");
    pageable = (float*)malloc((size_t)96 * 64 * 16);
    for (f = 0; f < frames; ++f) {
        p.u_time = (float)f / 60.0f;
        switch (f % 3) {
            case 0: st = sbx_multi_render_host(group, &p, (float*)pinned); break;
            case 1: st = sbx_multi_render_device(group, &p, &dev); if (st == SBX_OK) st = sbx_multi_sync(group); break;
            default: st = sbx_multi_render_host(group, &p, pageable); break;
        }
        if (st != SBX_OK) FAIL("frame %d: %s (%s)", f, sbx_strerror(st), sbx_multi_last_error(group));
        if (f % 97 == 0) nanosleep(&nap, NULL);
        if (f % 211 == 0 && sbx_multi_set_option(group, "record_events", (f / 211) & 1) != SBX_OK) FAIL("sbx_multi_set_option");
    }
    if (sbx_multi_last_timing(group, ms, 64) != SBX_OK) FAIL("sbx_multi_last_timing");
    sbx_host_free(sbx_multi_ctx(group, 0), pinned);
    sbx_multi_destroy(group);
    free(pageable);
    printf("{\"ok\": true, \"gpus\": %d, \"frames\": %d}\n", n, frames);
    return 0;
}
