/* abi_tour.c -- calls every entry point of include/sbx.h once or more with small, valid arguments (and a few invalid
 * ones).  Built by tests/test_host_mock_cpu.py with libsbx's sources and the mock driver (tests/native/fake_cuda.c) under
 * AddressSanitizer + UndefinedBehaviorSanitizer.  In the mock "device" memory is host memory of exactly the requested size,
 * so every copy the host library issues with a wrong size or offset (frame read-back, sequence staging, padded noise
 * volumes, IPC handles, flag words) is an ASan report here.  Nothing is rendered; pixels are checked on real GPUs.
 * prints one JSON line; exit code 0 = every call returned what it should */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sbx.h"

#define CHECK(call) do { int st_ = (call); if (st_ != SBX_OK) { printf("{\"ok\": false, \"call\": \"%s\", \"status\": %d, \"error\": \"%s\"}\n", #call, st_, sbx_last_error(ctx)); return 1; } } while (0)
#define EXPECT(call, want) do { int st_ = (call); if (st_ != (want)) { printf("{\"ok\": false, \"call\": \"%s\", \"status\": %d, \"wanted\": %d}\n", #call, st_, (int)(want)); return 1; } } while (0)

int main(void) {
    sbx_ctx* ctx = NULL;
    sbx_params p;
    sbx_shard sh = {4, 3, 1};
    sbx_frame_part part;
    sbx_timing tm;
    const int w = 100, h = 37;
    const size_t px = (size_t)w * h;
    float *dev = NULL, *dev2 = NULL, *imported = NULL, *alias = NULL;
    unsigned* flags = NULL;
    void* pinned = NULL;
    unsigned char handle[SBX_IPC_HANDLE_BYTES];
    unsigned char dds[148];
    float times[5] = {0.f, .1f, .2f, .3f, .4f};
    float* host = (float*)malloc(px * 16 * 5);
    unsigned char* host8 = (unsigned char*)malloc(px * 4);
    void* shared = NULL;
    int rows, k;
    static const char* apps[] = {"APP_EGG", "APP_CLOUDS", "APP_ATMOSPHERE", "APP_PLANET", "APP_RAYTRACER", "APP_SDF_AO", "APP_VINYL"};

    if (sbx_device_count() < 1) { printf("{\"ok\": false, \"error\": \"no device\"}\n"); return 3; }
    if (!sbx_version() || !sbx_strerror(SBX_ERR_NOMEM)) return 1;
    CHECK(sbx_default_params(&p, w, h));
    EXPECT(sbx_default_params(&p, 0, h), SBX_ERR_INVALID);
    CHECK(sbx_default_params(&p, w, h));
    rows = sbx_shard_rows(&sh, h);
    if (rows != 12) { printf("{\"ok\": false, \"error\": \"sbx_shard_rows = %d\"}\n", rows); return 1; }
    EXPECT(sbx_dds_volume_header(8, dds, 148), 148);

    CHECK(sbx_create(0, &ctx));
    EXPECT(sbx_render_host(ctx, &p, NULL, host), SBX_ERR_UNKNOWN_APP);      /* no app loaded yet */
    for (k = 0; k < 7; ++k) {
        CHECK(sbx_load_app(ctx, apps[k], NULL));
        CHECK(sbx_render_host(ctx, &p, NULL, host));
        CHECK(sbx_load_app(ctx, apps[k], "plugin"));
        CHECK(sbx_render_host(ctx, &p, &sh, host));
    }
    EXPECT(sbx_load_app(ctx, "APP_NOPE", NULL), SBX_ERR_UNKNOWN_APP);
    CHECK(sbx_load_app(ctx, "APP_CLOUDS", NULL));
    CHECK(sbx_last_timing(ctx, &tm));

    /* frames in device memory, every output flavour */
    CHECK(sbx_frame_alloc(ctx, px * 16 * 5, &dev));
    CHECK(sbx_frame_alloc(ctx, px * 16, &dev2));
    CHECK(sbx_render_device(ctx, &p, NULL, dev, NULL));
    CHECK(sbx_render_device(ctx, &p, &sh, dev, NULL));
    CHECK(sbx_unshard_device(ctx, w, h, &sh, dev, dev2, NULL));
    CHECK(sbx_render_device_rgba8(ctx, &p, NULL, (unsigned char*)dev, NULL));
    CHECK(sbx_render_host_rgba8(ctx, &p, NULL, host8));
    CHECK(sbx_render_host_rgba8(ctx, &p, &sh, host8));
    CHECK(sbx_render_sequence_device(ctx, &p, NULL, times, 5, dev, NULL));
    CHECK(sbx_render_sequence_host(ctx, &p, NULL, times, 5, host));
    CHECK(sbx_render_sequence_host(ctx, &p, &sh, times, 3, host));
    EXPECT(sbx_render_sequence_host(ctx, &p, NULL, times, 0, host), SBX_ERR_INVALID);
    CHECK(sbx_render_frame(ctx, &p, &sh, dev2, NULL));
    CHECK(sbx_frame_read(ctx, dev2, host, px * 16, NULL));
    sh.n_parts = -1;
    EXPECT(sbx_render_device(ctx, &p, &sh, dev, NULL), SBX_ERR_INVALID);
    sh.n_parts = 3;

    /* parts of a frame with completion flags in pinned memory */
    CHECK(sbx_host_alloc(ctx, 4096, (void**)&flags));
    memset(&part, 0, sizeof part);
    part.rows = sh; part.tile_parts = 2; part.tile_part = 1; part.done_flag = flags + 3; part.done_value = 9u;
    CHECK(sbx_render_frame_part(ctx, &p, &part, dev2, NULL));
    CHECK(sbx_stream_write_flag(ctx, flags, 9u, NULL));
    CHECK(sbx_stream_wait_flags(ctx, flags + 3, 1, 9u, NULL));
    part.tile_part = 2;
    EXPECT(sbx_render_frame_part(ctx, &p, &part, dev2, NULL), SBX_ERR_INVALID);

    /* host frames: pinned by the library, and caller memory registered for the device */
    CHECK(sbx_host_alloc(ctx, px * 16, &pinned));
    CHECK(sbx_render_host(ctx, &p, NULL, (float*)pinned));
    CHECK(sbx_last_timing(ctx, &tm));
    if (!tm.zero_copy) { printf("{\"ok\": false, \"error\": \"a pinned frame did not take the zero-copy path\"}\n"); return 1; }
    CHECK(sbx_render_host_rgba8(ctx, &p, NULL, (unsigned char*)pinned));
    if (posix_memalign(&shared, 4096, (px * 16 + 4095) / 4096 * 4096) != 0) return 1;
    CHECK(sbx_host_frame_register(ctx, shared, (px * 16 + 4095) / 4096 * 4096, &alias));
    CHECK(sbx_render_frame(ctx, &p, NULL, alias, NULL));
    CHECK(sbx_host_frame_unregister(ctx, shared));

    /* frames shared between processes */
    CHECK(sbx_frame_export(ctx, dev2, handle));
    CHECK(sbx_frame_import(ctx, handle, &imported));
    CHECK(sbx_render_frame(ctx, &p, &sh, imported, NULL));
    CHECK(sbx_frame_release(ctx, imported));

    /* options */
    CHECK(sbx_set_option(ctx, "use_hash_table", 0));
    CHECK(sbx_render_device(ctx, &p, NULL, dev, NULL));
    CHECK(sbx_set_option(ctx, "hash_table_log2", 12));
    CHECK(sbx_set_option(ctx, "use_hash_table", 1));
    CHECK(sbx_render_device(ctx, &p, NULL, dev, NULL));
    CHECK(sbx_set_option(ctx, "tail_waves_x100", 100));
    CHECK(sbx_render_device(ctx, &p, NULL, dev, NULL));
    CHECK(sbx_set_option(ctx, "tail_waves_x100", 0));
    CHECK(sbx_set_option(ctx, "record_events", 0));
    CHECK(sbx_render_frame(ctx, &p, NULL, dev2, NULL));
    EXPECT(sbx_set_option(ctx, "no_such_option", 1), SBX_ERR_INVALID);
    EXPECT(sbx_set_option(ctx, "hash_table_log2", 40), SBX_ERR_INVALID);
    CHECK(sbx_set_trace_buffer(ctx, (unsigned long long*)dev));
    CHECK(sbx_set_trace_buffer(ctx, NULL));

    /* the noise volume: bake, header, and the textured cloud path */
    {
        const int n = 8;
        float* vol = (float*)malloc((size_t)n * n * n * 16);
        float* dvol = NULL;
        float in[6] = {0.1f, 0.2f, 0.3f, 1.f, 2.f, 3.f}, out[2];
        CHECK(sbx_frame_alloc(ctx, (size_t)n * n * n * 16, &dvol));
        CHECK(sbx_bake_noise_volume_device(ctx, n, 0, n, dvol, NULL));
        CHECK(sbx_bake_noise_volume_host(ctx, n, 0, n, vol));
        CHECK(sbx_bake_noise_volume_host(ctx, n, 3, 2, vol));
        EXPECT(sbx_bake_noise_volume_host(ctx, n, 7, 2, vol), SBX_ERR_INVALID);
        CHECK(sbx_load_app(ctx, "APP_CLOUDS_TEX", NULL));
        EXPECT(sbx_render_host(ctx, &p, NULL, host), SBX_ERR_INVALID);       /* textures not set */
        CHECK(sbx_set_noise_volumes(ctx, vol, vol, n));
        CHECK(sbx_render_host(ctx, &p, NULL, host));
        CHECK(sbx_set_noise_volumes(ctx, vol, vol, n));                       /* replacing them */
        CHECK(sbx_load_app(ctx, "APP_CLOUDS_TEX", "tma"));
        CHECK(sbx_render_host(ctx, &p, &sh, host));
        CHECK(sbx_eval_op(ctx, "noise_iq", in, 3, out, 1, 2));
        EXPECT(sbx_eval_op(ctx, "no_such_op", in, 3, out, 1, 2), SBX_ERR_UNSUPPORTED);
        CHECK(sbx_frame_free(ctx, dvol));
        free(vol);
    }

    CHECK(sbx_host_free(ctx, pinned));
    CHECK(sbx_host_free(ctx, flags));
    CHECK(sbx_frame_free(ctx, dev));
    CHECK(sbx_frame_free(ctx, dev2));
    sbx_destroy(ctx);
    sbx_destroy(NULL);
    free(shared);
    free(host);
    free(host8);
    printf("{\"ok\": true}\n");
    return 0;
}
