// sbx_cli.cpp -- the C++ launcher: the stand-in for the reference's interactive hosts
// (VML SDL_app, src/Makefile:21; hlsltoy, util/hlsltoy/src/hlsltoy.cpp) that writes raw float4
// frames instead of presenting them.
//
//   sbx_cli compile <app_header.h> <APP_NAME> <out.cubin>      (needs no GPU)
//   sbx_cli render  <APP_NAME> <width> <height> <u_time> <out.rgba32f> [--variant native|plugin]
//                   [--steps N] [--frames K] [--device D] [--gpus N] [--ppm out.ppm] [--rgba8 out.rgba8]
//   sbx_cli flatten <file.h>                                   (needs no GPU)
//   sbx_cli bake <size> <out.dds> [--device D]                 (util/ddsvolgen: the 3-D noise texture, DDS + DX10 header + RGBA32F voxels)
// `render` prints one JSON line with the kernel time of the last frame.  --ppm / --rgba8 render the
// frame again through the 8-bit path (sbx_render_host_rgba8: what the reference's presenting hosts
// show, util/hlsltoy/src/hlsltoy.cpp:192) and write it as a binary PPM (top row first) / raw bytes.
// `flatten` is the reference's include expander (util/inclxpnd/src/inclxpnd.cpp:8-41): prints the
// file with every `#include "x"` / `#include <x>` line replaced, recursively, by the contents of x
// (resolved against the current directory, as the reference does); a missing file prints
// "*** error: cannot include file: x" in its place.  It is how an app header is made pasteable
// into shadertoy (README.md:29-32).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/sbx.h"

// util/inclxpnd/src/inclxpnd.cpp:8-41 -- same tokenisation (first whitespace-separated token == "#include",
// second token quoted with "" or <>), same error codes (2: token too short, 3: not quoted)
static int flatten(std::istream& in, std::ostream& out) {
    for (std::string line; std::getline(in, line);) {
        std::istringstream tok(line);
        std::string word;
        tok >> word;
        if (word != "#include") { out << line << std::endl; continue; }
        word.clear();
        tok >> word;
        if (word.length() < 2) return 2;
        if ((word.front() != '"' && word.front() != '<') || (word.back() != '"' && word.back() != '>')) return 3;
        const std::string name(word.begin() + 1, word.end() - 1);
        std::ifstream inc(name);
        if (inc.good()) flatten(inc, out);
        else out << "*** error: cannot include file: " << name << std::endl;
    }
    return 0;
}

static int usage() {
    fprintf(stderr,
            "usage: sbx_cli compile <app_header.h> <APP_NAME> <out.cubin>\n"
            "       sbx_cli render <APP_NAME> <width> <height> <u_time> <out.rgba32f|-> [--variant v] [--steps n] "
            "[--frames k] [--device d] [--gpus n] [--ppm out.ppm] [--rgba8 out.rgba8]\n"
            "       sbx_cli flatten <file.h>\n"
            "       sbx_cli bake <size> <out.dds> [--device d]\n");
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
    if (!strcmp(argv[1], "flatten")) {
        if (argc < 3) return usage();
        std::ifstream in(argv[2]);
        if (!in.good()) return 1;
        return flatten(in, std::cout);
    }
    if (!strcmp(argv[1], "bake")) {   // util/ddsvolgen/src/ddsvolgen.cpp:64-149 with the voxel loop on the GPU
        if (argc < 4) return usage();
        const int size = atoi(argv[2]);
        int device = 0;
        for (int i = 4; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
            else return usage();
        }
        unsigned char header[148];
        if (size <= 0 || sbx_dds_volume_header(size, header, sizeof header) != 148) return usage();
        sbx_ctx* ctx = nullptr;
        int st = sbx_create(device, &ctx);
        if (st != SBX_OK) { fprintf(stderr, "sbx_create: %s: %s\n", sbx_strerror(st), sbx_last_error(nullptr)); return 1; }
        std::vector<float> vol((size_t)size * size * size * 4);
        st = sbx_bake_noise_volume_host(ctx, size, 0, size, vol.data());
        if (st != SBX_OK) { fprintf(stderr, "sbx_bake_noise_volume_host: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
        sbx_timing tm{};
        sbx_last_timing(ctx, &tm);
        FILE* fp = fopen(argv[3], "wb");
        if (!fp) { perror(argv[3]); return 2; }
        fwrite(header, 1, sizeof header, fp);
        fwrite(vol.data(), sizeof(float), vol.size(), fp);
        fclose(fp);
        printf("{\"baked\": \"%s\", \"size\": %d, \"kernel_ms\": %.3f, \"mvoxel_per_s\": %.1f}\n", argv[3], size, tm.kernel_ms,
               tm.kernel_ms > 0 ? (double)size * size * size / tm.kernel_ms * 1e-3 : 0.0);
        sbx_destroy(ctx);
        return 0;
    }
    if (!strcmp(argv[1], "render")) {
        if (argc < 7) return usage();
        const char* app = argv[2];
        const int w = atoi(argv[3]), h = atoi(argv[4]);
        const float t = (float)atof(argv[5]);
        const char* out_path = argv[6];
        const char* variant = nullptr;
        const char* ppm_path = nullptr;
        const char* rgba8_path = nullptr;
        int steps = 0, frames = 1, device = 0, gpus = 1;
        for (int i = 7; i + 1 < argc; i += 2) {
            if (!strcmp(argv[i], "--variant")) variant = argv[i + 1];
            else if (!strcmp(argv[i], "--steps")) steps = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--frames")) frames = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--gpus")) gpus = atoi(argv[i + 1]);
            else if (!strcmp(argv[i], "--ppm")) ppm_path = argv[i + 1];
            else if (!strcmp(argv[i], "--rgba8")) rgba8_path = argv[i + 1];
            else return usage();
        }
        sbx_params p;
        if (w <= 0 || h <= 0 || sbx_default_params(&p, w, h) != SBX_OK) { fprintf(stderr, "render: bad frame size %dx%d\n", w, h); return 2; }
        p.u_time = t;
        if (steps > 0) p.cld_march_steps = steps;
        // one GPU: a context; several: a group (one process, one part of every frame per GPU, sbx_multi_*)
        sbx_multi* group = nullptr;
        sbx_ctx* ctx = nullptr;
        int st;
        if (gpus > 1) {
            // parts beyond the GPUs the box has share them round-robin (the N-part path on a smaller box)
            const int have = sbx_device_count();
            std::vector<int> devices;
            for (int i = 0; i < gpus; ++i) devices.push_back(have > 0 ? (device + i) % have : i);
            st = sbx_multi_create(devices.data(), gpus, &group);
            if (st != SBX_OK) { fprintf(stderr, "sbx_multi_create(%d): %s: %s\n", gpus, sbx_strerror(st), sbx_last_error(nullptr)); return 1; }
            st = sbx_multi_load_app(group, app, variant);
            if (st != SBX_OK) { fprintf(stderr, "sbx_multi_load_app: %s: %s\n", sbx_strerror(st), sbx_multi_last_error(group)); return 1; }
            ctx = sbx_multi_ctx(group, 0);
        } else {
            st = sbx_create(device, &ctx);
            if (st != SBX_OK) { fprintf(stderr, "sbx_create: %s: %s\n", sbx_strerror(st), sbx_last_error(nullptr)); return 1; }
            st = sbx_load_app(ctx, app, variant);
            if (st != SBX_OK) { fprintf(stderr, "sbx_load_app: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
        }
        // the frame lives in pinned + mapped host memory: the render kernels store into it directly
        void* frame_mem = nullptr;
        const size_t frame_floats = (size_t)w * h * 4;
        st = sbx_host_alloc(ctx, frame_floats * sizeof(float), &frame_mem);
        if (st != SBX_OK) { fprintf(stderr, "sbx_host_alloc: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
        float* frame = (float*)frame_mem;
        sbx_timing tm{};
        for (int f = 0; f < frames; ++f) {
            st = group ? sbx_multi_render_host(group, &p, frame) : sbx_render_host(ctx, &p, nullptr, frame);
            if (st != SBX_OK) { fprintf(stderr, "render: %s: %s\n", sbx_strerror(st), group ? sbx_multi_last_error(group) : sbx_last_error(ctx)); return 1; }
            sbx_last_timing(ctx, &tm);
        }
        if (strcmp(out_path, "-")) {
            FILE* fp = fopen(out_path, "wb");
            if (!fp) { perror(out_path); return 1; }
            fwrite(frame, sizeof(float), frame_floats, fp);
            fclose(fp);
        }
        if (ppm_path || rgba8_path) {
            std::vector<unsigned char> px((size_t)w * h * 4);
            st = sbx_render_host_rgba8(ctx, &p, nullptr, px.data());
            if (st != SBX_OK) { fprintf(stderr, "sbx_render_host_rgba8: %s: %s\n", sbx_strerror(st), sbx_last_error(ctx)); return 1; }
            if (rgba8_path) {
                FILE* fp = fopen(rgba8_path, "wb");
                if (!fp) { perror(rgba8_path); return 1; }
                fwrite(px.data(), 1, px.size(), fp);
                fclose(fp);
            }
            if (ppm_path) {   // P6, top row first: frame row 0 is the BOTTOM row (fragCoord.y = 0.5)
                FILE* fp = fopen(ppm_path, "wb");
                if (!fp) { perror(ppm_path); return 1; }
                fprintf(fp, "P6\n%d %d\n255\n", w, h);
                std::vector<unsigned char> row((size_t)w * 3);
                for (int y = h - 1; y >= 0; --y) {
                    const unsigned char* src = px.data() + (size_t)y * w * 4;
                    for (int x = 0; x < w; ++x) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
                    fwrite(row.data(), 1, row.size(), fp);
                }
                fclose(fp);
            }
        }
        printf("{\"app\": \"%s\", \"width\": %d, \"height\": %d, \"u_time\": %g, \"gpus\": %d, \"kernel_ms\": %.4f, \"d2h_ms\": %.4f, "
               "\"mpix_per_s\": %.2f, \"grid\": %d, \"block\": %d, \"regs\": %d, \"ctas_per_sm\": %d}\n",
               app, w, h, t, gpus > 1 ? gpus : 1, tm.kernel_ms, tm.d2h_ms, tm.kernel_ms > 0 ? (double)w * h / tm.kernel_ms * 1e-3 : 0.0,
               tm.grid_blocks, tm.block_threads, tm.regs_per_thread, tm.blocks_per_sm);
        sbx_host_free(ctx, frame_mem);
        if (group) sbx_multi_destroy(group); else sbx_destroy(ctx);
        return 0;
    }
    return usage();
}
