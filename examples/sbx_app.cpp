// sbx_app.cpp -- the C++ host of INTEGRATION.md §2: what replaces VML's SDL_app.cpp (src/Makefile:15,21-22) when a
// shaderbox app is rendered on a B200.  The app header is read as it is; this program never touches it.
//
//   g++ -std=c++17 -O2 examples/sbx_app.cpp -Iinclude -Lshaderbox_b200 -lsbx -Wl,-rpath,$PWD/shaderbox_b200 -o sbx_app
//   ./sbx_app /path/to/shaderbox/src/app_planet.h APP_PLANET 1920 1080 120 frames/planet_%04d.ppm
// The frame lives in memory from sbx_host_alloc (pinned + mapped): the render kernel stores its pixels straight into
// it over PCIe, so there is no device->host copy to wait for (a std::vector frame works too, at about half the rate:
// rendered in HBM, then copied -- bench.py reports both as e2e.value / e2e.pageable_value).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "sbx.h"

static int fail(sbx_ctx* ctx, const char* what, int st) {
    std::fprintf(stderr, "%s: %s (%s)\n", what, sbx_strerror(st), sbx_last_error(ctx));
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 7) {
        std::fprintf(stderr, "usage: %s <app_header.h> <APP_NAME> <width> <height> <frames> <out_%%04d.ppm>\n", argv[0]);
        return 2;
    }
    const char* header = argv[1];
    const char* app = argv[2];
    const int w = std::atoi(argv[3]), h = std::atoi(argv[4]), frames = std::atoi(argv[5]);

    sbx_ctx* ctx = nullptr;
    int st = sbx_create(/*device*/ 0, &ctx);                       // fails loudly without a B200: there is no CPU path
    if (st != SBX_OK) return fail(nullptr, "sbx_create", st);
    st = sbx_compile_app(ctx, header, app, nullptr);               // == $(CXX) $(APP) -c SDL_app.cpp, for sm_100a
    if (st != SBX_OK) return fail(ctx, "sbx_compile_app", st);
    st = sbx_load_app(ctx, app, "plugin");                         // == APP = -DAPP_PLANET
    if (st != SBX_OK) return fail(ctx, "sbx_load_app", st);

    sbx_params p;
    sbx_default_params(&p, w, h);                                  // src/uniform_buffer.h defaults
    void* frame_mem = nullptr;                                     // the host frame: pinned + mapped, the kernel writes into it
    st = sbx_host_alloc(ctx, size_t(w) * h * 4, &frame_mem);
    if (st != SBX_OK) return fail(ctx, "sbx_host_alloc", st);
    unsigned char* rgba8 = static_cast<unsigned char*>(frame_mem);
    std::vector<unsigned char> row(size_t(w) * 3);
    for (int f = 0; f < frames; ++f) {
        p.u_time = f / 60.0f;                                      // iGlobalTime of a 60 Hz host
        st = sbx_render_host_rgba8(ctx, &p, nullptr, rgba8);          // the frame as the reference's swap chain holds it
        if (st != SBX_OK) return fail(ctx, "sbx_render_host_rgba8", st);
        char name[512];
        std::snprintf(name, sizeof name, argv[6], f);
        if (FILE* fp = std::fopen(name, "wb")) {                   // "present": a PPM, top row first
            std::fprintf(fp, "P6\n%d %d\n255\n", w, h);
            for (int y = h - 1; y >= 0; --y) {
                for (int x = 0; x < w; ++x)
                    for (int c = 0; c < 3; ++c) row[size_t(3) * x + c] = rgba8[(size_t(y) * w + x) * 4 + c];
                std::fwrite(row.data(), 1, row.size(), fp);
            }
            std::fclose(fp);
        }
    }
    sbx_timing tm{};
    sbx_last_timing(ctx, &tm);
    std::printf("{\"app\": \"%s\", \"frames\": %d, \"last_kernel_ms\": %.3f}\n", app, frames, tm.kernel_ms);
    sbx_host_free(ctx, frame_mem);
    sbx_destroy(ctx);
    return 0;
}
