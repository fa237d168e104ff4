/* multi_c.c -- a plain C99 host that renders one frame over N parts through libsbx.so (sbx_multi_*), the way the
 * reference's single-process hosts would (util/hlsltoy/src/hlsltoy.cpp:494-520 renders every frame from one loop),
 * and checks it byte for byte against the 1-GPU frame of sbx_render_host.
 * usage: multi_c <n_parts> [APP] [width] [height]     parts beyond the GPUs present share them round-robin
 * prints one JSON line; exit code 0 = frames identical, 3 = no device (the caller decides whether that is a failure) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sbx.h"

#define FAIL(...) do { printf("{\"ok\": false, \"error\": \""); printf(__VA_ARGS__); printf("\"}\n"); return 1; } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2;
    const char* app = argc > 2 ? argv[2] : "APP_CLOUDS";
    const int w = argc > 3 ? atoi(argv[3]) : 333, h = argc > 4 ? atoi(argv[4]) : 187;
    const int have = sbx_device_count();
    int devices[64], i, st, frames;
    sbx_params p;
    sbx_ctx* one = NULL;
    sbx_multi* group = NULL;
    float *want, *pageable;
    void* pinned = NULL;
    float ms[64];
    size_t bytes;

    if (have <= 0) { printf("{\"ok\": false, \"error\": \"no CUDA device\"}\n"); return 3; }
    if (n < 1 || n > 64 || sbx_default_params(&p, w, h) != SBX_OK) FAIL("bad arguments");
    p.u_time = 1.5f;
    p.cld_march_steps = 64;
    bytes = (size_t)w * h * 4 * sizeof(float);
    want = (float*)malloc(bytes);
    pageable = (float*)malloc(bytes);
    if (!want || !pageable) FAIL("out of memory");

    /* the 1-GPU frame */
    if ((st = sbx_create(0, &one)) != SBX_OK) FAIL("sbx_create: %s", sbx_strerror(st));
    if ((st = sbx_load_app(one, app, NULL)) != SBX_OK) FAIL("sbx_load_app: %s", sbx_last_error(one));
    if ((st = sbx_render_host(one, &p, NULL, want)) != SBX_OK) FAIL("sbx_render_host: %s", sbx_last_error(one));

    /* the same frame over n parts from this process */
    for (i = 0; i < n; ++i) devices[i] = i % have;
    if ((st = sbx_multi_create(devices, n, &group)) != SBX_OK) FAIL("sbx_multi_create: %s", sbx_strerror(st));
    if (sbx_multi_gpus(group) != n) FAIL("sbx_multi_gpus");
    if ((st = sbx_multi_load_app(group, app, NULL)) != SBX_OK) FAIL("sbx_multi_load_app: %s", sbx_multi_last_error(group));
    memset(pageable, 0xff, bytes);
    if ((st = sbx_multi_render_host(group, &p, pageable)) != SBX_OK) FAIL("sbx_multi_render_host: %s", sbx_multi_last_error(group));
    if (memcmp(pageable, want, bytes) != 0) FAIL("pageable frame differs from the 1-GPU frame");

    /* a pinned frame from sbx_host_alloc: every part stores into it directly; several frames, u_time advancing */
    if ((st = sbx_host_alloc(sbx_multi_ctx(group, 0), bytes, &pinned)) != SBX_OK) FAIL("sbx_host_alloc: %s", sbx_strerror(st));
    for (frames = 0; frames < 3; ++frames) {
        p.u_time = 1.5f + (float)frames;
        if ((st = sbx_render_host(one, &p, NULL, want)) != SBX_OK) FAIL("sbx_render_host");
        memset(pinned, 0xff, bytes);
        if ((st = sbx_multi_render_host(group, &p, (float*)pinned)) != SBX_OK) FAIL("sbx_multi_render_host (pinned): %s", sbx_multi_last_error(group));
        if (memcmp(pinned, want, bytes) != 0) FAIL("pinned frame %d differs from the 1-GPU frame", frames);
    }
    if (sbx_multi_last_timing(group, ms, 64) != SBX_OK) FAIL("sbx_multi_last_timing");
    for (i = 0; i < n; ++i) if (!(ms[i] > 0.0f)) FAIL("part %d reports no kernel time", i);

    /* errors stay errors */
    if (sbx_multi_render_host(group, &p, NULL) != SBX_ERR_INVALID) FAIL("NULL frame accepted");
    if (sbx_multi_load_app(group, "APP_NOPE", NULL) != SBX_ERR_UNKNOWN_APP) FAIL("unknown app accepted");

    printf("{\"ok\": true, \"app\": \"%s\", \"parts\": %d, \"gpus\": %d, \"width\": %d, \"height\": %d, \"kernel_ms_part0\": %.4f}\n",
           app, n, have < n ? have : n, w, h, ms[0]);
    sbx_host_free(sbx_multi_ctx(group, 0), pinned);
    sbx_multi_destroy(group);
    sbx_destroy(one);
    free(want);
    free(pageable);
    return 0;
}
