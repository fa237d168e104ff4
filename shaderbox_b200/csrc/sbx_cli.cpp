// sbx_cli.cpp -- the C++ launcher: the stand-in for the reference's interactive hosts
// (VML SDL_app, src/Makefile:21; hlsltoy, util/hlsltoy/src/hlsltoy.cpp) that writes raw float4
// frames instead of presenting them.
//
//   sbx_cli compile <app_header.h> <APP_NAME> <out.cubin>      (needs no GPU)
//   sbx_cli render  <APP_NAME> <width> <height> <u_time> <out.rgba32f> [--variant native|plugin]
//                   [--steps N] [--frames K] [--device D]
// `render` prints one JSON line with the kernel time of the last frame.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sbx.h"

static int usage() {
    fprintf(stderr,
            "usage: sbx_cli compile <app_header.h> <APP_NAME> <out.cubin>\n"
            "       sbx_cli render <APP_NAME> <width> <height> <u_time> <out.rgba32f|-> [--variant v] [--steps n] "
            "[--frames k] [--device d]\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    if (!strcmp(argv[1], "compile")) {
        if (argc < 5) return usage();
        const int st = sbx_compile_app(nullptr, argv[2], argv[3], argv[4]);
        const char* log = sbx_last_error(nullptr);
        if (log && *log) fprintf(stderr, "%s\n", log);
        if (st != SBX_OK) { fprintf(stderr, "compile failed: %s\n", sbx_strerror(st)); return 1; }
        printf("{\"compiled\": \"%s\", \"app\": \"%s\", \"image\": \"%s\"}\n", argv[2], argv[3], argv[4]);
        return 0;
    }
    if (!strcmp(argv[1], "render")) {
        if (argc < 7) return usage();
        const char* app = argv[2];
        const int w = atoi(argv[3]), h = atoi(argv[4]);
        const float t = (float)atof(argv[5]);
        const char* out_path = argv[6];
        const char* variant = nullptr;
        int steps = 0, frames = 1, device = 0;
        for (int i = 7; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--variant")) variant = argv[i + 1];
            else if (!strcmp(argv[i], "--steps")) steps = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--frames")) frames = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
            else return usage();
        }
        sbx_ctx* ctx = nullptr;
        int st = sbx_create(device, &ctx);
        if (st != SBX_OK) { fprintf(stderr, "sbx_create: %s: %s\n", sbx_strerror(st), sbx_last_error(nullptr)); return 1; }
        st = sbx_load_app(ctx, app, variant);
        if (st != SBX_OK) { fprintf(stderr, "sbx_load_app: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
        sbx_params p;
        sbx_default_params(&p, w, h);
        p.u_time = t;
        if (steps > 0) p.cld_march_steps = steps;
        std::vector<float> frame((size_t)w * h * 4);
        sbx_timing tm{};
        for (int f = 0; f < frames; ++f) {
            st = sbx_render_host(ctx, &p, nullptr, frame.data());
            if (st != SBX_OK) { fprintf(stderr, "sbx_render_host: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
            sbx_last_timing(ctx, &tm);
        }
        if (strcmp(out_path, "-")) {
            FILE* fp = fopen(out_path, "wb");
            if (!fp) { perror(out_path); return 1; }
            fwrite(frame.data(), sizeof(float), frame.size(), fp);
            fclose(fp);
        }
        printf("{\"app\": \"%s\", \"width\": %d, \"height\": %d, \"u_time\": %g, \"kernel_ms\": %.4f, \"d2h_ms\": %.4f, "
               "\"mpix_per_s\": %.2f, \"grid\": %d, \"block\": %d, \"regs\": %d, \"ctas_per_sm\": %d}\n",
               app, w, h, t, tm.kernel_ms, tm.d2h_ms, tm.kernel_ms > 0 ? (double)w * h / tm.kernel_ms * 1e-3 : 0.0,
               tm.grid_blocks, tm.block_threads, tm.regs_per_thread, tm.blocks_per_sm);
        sbx_destroy(ctx);
        return 0;
    }
    return usage();
}
