// sbx_host.cpp -- the C ABI of include/sbx.h: a host for shaderbox app shaders on one B200.
//
// Plain C++ over the CUDA driver API (loaded lazily with dlopen so the library itself loads, and
// its symbols can be checked, on a machine without a driver).  Kernels come as sm_100a cubins:
//   images/<APP>.plugin.cubin   an UNCHANGED shaderbox app header compiled by sbx_compile_app
//   images/<APP>.native.cubin   hand-written scene kernels (when present)
//   images/sbx_util.cubin       hash-table / unshard / operator-evaluation kernels
// There is deliberately no CPU path: without a device every render entry fails with an error.
#include <cuda.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "../include/sbx/sbx_launch.h"
#include "../include/sbx/sbx_math.h"
#include "sbx_internal.h"

namespace {

using sbx::driver_api;

std::mutex g_mutex;
std::string g_last_error;   // for calls without a context

}  // namespace

sbx::driver_api* sbx::load_driver(std::string* err) {
    static driver_api api;
    std::lock_guard<std::mutex> lock(g_mutex);
    if (api.lib) return &api;
    void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        *err = "no CUDA driver (libcuda.so.1 not found): shaderbox_b200 has no CPU path";
        return nullptr;
    }
#define SBX_SYM(field, name)                                                            \
    *(void**)(&api.field) = dlsym(lib, name);                                           \
    if (!api.field) { *err = std::string("libcuda lacks ") + name; return nullptr; }
    SBX_SYM(Init, "cuInit")
    SBX_SYM(DeviceGet, "cuDeviceGet")
    SBX_SYM(DeviceGetAttribute, "cuDeviceGetAttribute")
    SBX_SYM(DevicePrimaryCtxRetain, "cuDevicePrimaryCtxRetain")
    SBX_SYM(DevicePrimaryCtxRelease, "cuDevicePrimaryCtxRelease_v2")
    SBX_SYM(CtxPushCurrent, "cuCtxPushCurrent_v2")
    SBX_SYM(CtxPopCurrent, "cuCtxPopCurrent_v2")
    SBX_SYM(CtxSynchronize, "cuCtxSynchronize")
    SBX_SYM(ModuleLoadData, "cuModuleLoadData")
    SBX_SYM(ModuleUnload, "cuModuleUnload")
    SBX_SYM(ModuleGetFunction, "cuModuleGetFunction")
    SBX_SYM(ModuleGetGlobal, "cuModuleGetGlobal_v2")
    SBX_SYM(FuncGetAttribute, "cuFuncGetAttribute")
    SBX_SYM(FuncSetAttribute, "cuFuncSetAttribute")
    SBX_SYM(OccupancyMaxActiveBlocksPerMultiprocessor, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
    SBX_SYM(MemAlloc, "cuMemAlloc_v2")
    SBX_SYM(MemFree, "cuMemFree_v2")
    SBX_SYM(MemcpyHtoD, "cuMemcpyHtoD_v2")
    SBX_SYM(MemcpyDtoH, "cuMemcpyDtoH_v2")
    SBX_SYM(MemcpyDtoHAsync, "cuMemcpyDtoHAsync_v2")
    SBX_SYM(MemcpyHtoDAsync, "cuMemcpyHtoDAsync_v2")
    SBX_SYM(LaunchKernel, "cuLaunchKernel")
    SBX_SYM(StreamSynchronize, "cuStreamSynchronize")
    SBX_SYM(EventCreate, "cuEventCreate")
    SBX_SYM(EventRecord, "cuEventRecord")
    SBX_SYM(EventSynchronize, "cuEventSynchronize")
    SBX_SYM(EventElapsedTime, "cuEventElapsedTime")
    SBX_SYM(EventDestroy, "cuEventDestroy_v2")
    SBX_SYM(GetErrorString, "cuGetErrorString")
    SBX_SYM(PointerGetAttribute, "cuPointerGetAttribute")
    SBX_SYM(MemHostRegister, "cuMemHostRegister_v2")
    SBX_SYM(MemHostUnregister, "cuMemHostUnregister")
    SBX_SYM(MemHostGetDevicePointer, "cuMemHostGetDevicePointer_v2")
    SBX_SYM(IpcGetMemHandle, "cuIpcGetMemHandle")
    SBX_SYM(IpcOpenMemHandle, "cuIpcOpenMemHandle_v2")
    SBX_SYM(IpcCloseMemHandle, "cuIpcCloseMemHandle")
    SBX_SYM(MemsetD32, "cuMemsetD32_v2")
    SBX_SYM(MemHostAlloc, "cuMemHostAlloc")
    SBX_SYM(MemFreeHost, "cuMemFreeHost")
    SBX_SYM(StreamWaitValue32, "cuStreamWaitValue32_v2")
    SBX_SYM(StreamWriteValue32, "cuStreamWriteValue32_v2")
    SBX_SYM(StreamBatchMemOp, "cuStreamBatchMemOp_v2")
    SBX_SYM(TensorMapEncodeTiled, "cuTensorMapEncodeTiled")
    SBX_SYM(StreamWaitEvent, "cuStreamWaitEvent")
    SBX_SYM(StreamCreate, "cuStreamCreate")
    SBX_SYM(StreamDestroy, "cuStreamDestroy_v2")
    SBX_SYM(MemAllocAsync, "cuMemAllocAsync")
    SBX_SYM(MemFreeAsync, "cuMemFreeAsync")
    SBX_SYM(DeviceGetCount, "cuDeviceGetCount")
    SBX_SYM(DeviceCanAccessPeer, "cuDeviceCanAccessPeer")
    SBX_SYM(CtxEnablePeerAccess, "cuCtxEnablePeerAccess")
#undef SBX_SYM
    api.lib = lib;
    return &api;
}

namespace {
using sbx::load_driver;

struct kernel_image {
    CUmodule module = nullptr;
    CUfunction render = nullptr;
    int regs = 0, max_threads = 0, blocks_per_sm = 0;
    int warps_per_cta = 4;
    int tile_w = SBX_TILE_W, tile_h = SBX_TILE_H, lanes_per_pixel = 1;   // sbx_image_info of the image
    int hybrid_lanes = 0;         // > 1: the image has a second region marched with this many lanes per pixel
    int uses_noise_tex = 0;          // sbx_image_hints[1]: sbx_render takes a second parameter (sbx_tex_params)
    int trivial_rows_permille = 0;   // sbx_image_hints[0]: this share of the frame's bottom rows is trivial (issued last)
    std::string variant;
};

bool read_file(const std::string& path, std::string* out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    *out = ss.str();
    return true;
}

}  // namespace

struct sbx_ctx {
    driver_api* cu = nullptr;
    int device = 0;
    CUdevice dev = 0;
    CUcontext ctx = nullptr;
    int sm_count = 0;
    std::string last_error;
    sbx_timing timing{};

    std::map<std::string, kernel_image> images;   // "APP_X/variant"
    kernel_image* current = nullptr;
    kernel_image* current_coop = nullptr;   // cooperative images of the same app (4 and 2 lanes per pixel), for small grids
    kernel_image* current_coop2 = nullptr;
    kernel_image* current_hybrid = nullptr; // one lane per pixel first, the tail of the launch with 4 lanes per pixel
    std::string current_app;

    CUmodule util_module = nullptr;
    CUfunction k_hash = nullptr, k_unshard = nullptr, k_eval = nullptr, k_bake = nullptr;

    CUdeviceptr lut = 0;          // SBX_LUT_MATH_BYTES
    CUdeviceptr hash_tab = 0;     // 2 float4 per entry (noise_iq.h)
    int hash_lo = 0, hash_len = 0;
    int opt_hash_log2 = 18;       // table covers [-2^(k-1), 2^(k-1))
    int opt_use_hash = 1;
    int opt_zero_copy = 1;        // sbx_render_host: store straight into pinned+mapped host frames
    int opt_coop_waves_x100 = 250;   // use the cooperative image when the grid is below this many waves of resident warps
    int opt_trivial_rows_last = 1;   // honour sbx_image_hints[0]
    int opt_record_events = 1;       // bracket every render launch with timing events (sbx_last_timing.kernel_ms)
    int opt_tail_waves_x100 = 0;     // hybrid image (opt-in): march the last this-many waves of the launch with 4 lanes per pixel
    int opt_tail_max_waves_x100 = 1200;  // ... for launches below this many waves (a long launch amortises its tail anyway)
    CUdeviceptr noise_vol[2] = {0, 0};   // padded single-channel noise textures (sbx_set_noise_volumes)
    sbx_tex_params tex{};
    CUdeviceptr trace = 0;           // profiling hook: per-warp records of the next launches (trace images only)
    CUdeviceptr done_counters = 0;   // ring of CTA counters for launches that signal a completion flag
    unsigned done_seq = 0;

    // u_time values of sequence launches travel through a small ring of pinned staging slots (so the copy is truly
    // asynchronous and the caller's array is consumed before the call returns); the device copy is stream-ordered
    static const int kTimeSlots = 8, kTimeSlotFloats = 65536;
    float* times_ring = nullptr;  // kTimeSlots * kTimeSlotFloats pinned floats
    CUevent times_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned times_seq = 0;
    CUdeviceptr frame = 0;        // internal frame for sbx_render_host
    size_t frame_bytes = 0;
    CUevent ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;

    int fail(int status, const char* fmt, ...) {
        char buf[2048];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        last_error = buf;
        return status;
    }
    int check(CUresult r, const char* what) {
        if (r == CUDA_SUCCESS) return SBX_OK;
        const char* s = nullptr;
        cu->GetErrorString(r, &s);
        return fail(SBX_ERR_CUDA, "%s: %s (%d)", what, s ? s : "?", (int)r);
    }
};

CUcontext sbx::context_of(sbx_ctx* ctx) { return ctx ? ctx->ctx : nullptr; }

namespace {

struct ctx_scope {   // make the primary context current for the duration of a call
    sbx_ctx* c;
    explicit ctx_scope(sbx_ctx* c_) : c(c_) { c->cu->CtxPushCurrent(c->ctx); }
    ~ctx_scope() { CUcontext old; c->cu->CtxPopCurrent(&old); }
};

#define SBX_TRY(expr, what)                                     \
    do {                                                        \
        int st_ = ctx->check((expr), what);                     \
        if (st_ != SBX_OK) return st_;                          \
    } while (0)

int load_util(sbx_ctx* ctx) {
    if (ctx->util_module) return SBX_OK;
    std::string bin;
    const std::string path = sbx::library_dir() + "/images/sbx_util.cubin";
    if (!read_file(path, &bin)) return ctx->fail(SBX_ERR_UNKNOWN_APP, "kernel image %s not found (run __graft_entry__.build())", path.c_str());
    SBX_TRY(ctx->cu->ModuleLoadData(&ctx->util_module, bin.data()), "cuModuleLoadData(sbx_util)");
    SBX_TRY(ctx->cu->ModuleGetFunction(&ctx->k_hash, ctx->util_module, "sbx_hash_table_kernel"), "get sbx_hash_table_kernel");
    SBX_TRY(ctx->cu->ModuleGetFunction(&ctx->k_unshard, ctx->util_module, "sbx_unshard_kernel"), "get sbx_unshard_kernel");
    SBX_TRY(ctx->cu->ModuleGetFunction(&ctx->k_eval, ctx->util_module, "sbx_eval_op_kernel"), "get sbx_eval_op_kernel");
    SBX_TRY(ctx->cu->ModuleGetFunction(&ctx->k_bake, ctx->util_module, "sbx_bake_volume_kernel"), "get sbx_bake_volume_kernel");
    return SBX_OK;
}

// LUT block + lattice-hash memo; (re)built when the table option changes
int ensure_tables(sbx_ctx* ctx, CUstream stream, bool force_table = false) {
    int st = load_util(ctx);
    if (st != SBX_OK) return st;
    if (!ctx->lut) {
        static const sbx_u64 exp2_tab[32] = {SBX_EXP2_TABLE_INIT};
        static const double log2_tab[32] = {SBX_LOG2_TABLE_INIT};
        unsigned char block[SBX_LUT_MATH_BYTES];
        std::memcpy(block, exp2_tab, sizeof exp2_tab);
        std::memcpy(block + sizeof exp2_tab, log2_tab, sizeof log2_tab);
        SBX_TRY(ctx->cu->MemAlloc(&ctx->lut, SBX_LUT_MATH_BYTES), "cuMemAlloc(lut)");
        SBX_TRY(ctx->cu->MemcpyHtoD(ctx->lut, block, SBX_LUT_MATH_BYTES), "cuMemcpyHtoD(lut)");
    }
    // hand-written kernels index the memo table unconditionally (misses are detected, not avoided)
    const int want_len = (ctx->opt_use_hash || force_table) ? (1 << ctx->opt_hash_log2) : 0;
    if (want_len != ctx->hash_len) {
        if (ctx->hash_tab) { ctx->cu->StreamSynchronize(stream); ctx->cu->MemFree(ctx->hash_tab); ctx->hash_tab = 0; }
        ctx->hash_len = 0;
        if (want_len > 0) {
            SBX_TRY(ctx->cu->MemAlloc(&ctx->hash_tab, (size_t)want_len * 8 * sizeof(float)), "cuMemAlloc(hash table)");
            int lo = -(want_len / 2), len = want_len;
            void* args[] = {&ctx->hash_tab, &lo, &len, &ctx->lut};
            SBX_TRY(ctx->cu->LaunchKernel(ctx->k_hash, (unsigned)((len + 255) / 256), 1, 1, 256, 1, 1,
                                          SBX_LUT_MATH_BYTES, stream, args, nullptr),
                    "launch sbx_hash_table_kernel");
            // the table is per-context state shared by every later launch on ANY stream: finish it here, once
            SBX_TRY(ctx->cu->StreamSynchronize(stream), "cuStreamSynchronize(hash table)");
            ctx->hash_lo = lo;
            ctx->hash_len = len;
            ctx->timing.launches += 1;
        }
    }
    return SBX_OK;
}

int bind_image(sbx_ctx* ctx, const std::string& key, const std::string& cubin, const std::string& variant) {
    kernel_image img;
    SBX_TRY(ctx->cu->ModuleLoadData(&img.module, cubin.data()), "cuModuleLoadData(app image)");
    SBX_TRY(ctx->cu->ModuleGetFunction(&img.render, img.module, "sbx_render"), "cuModuleGetFunction(sbx_render)");
    ctx->cu->FuncGetAttribute(&img.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, img.render);
    ctx->cu->FuncGetAttribute(&img.max_threads, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, img.render);
    img.warps_per_cta = img.max_threads >= 32 ? img.max_threads / 32 : 4;   // __launch_bounds__ = CTA size
    ctx->cu->OccupancyMaxActiveBlocksPerMultiprocessor(&img.blocks_per_sm, img.render, img.warps_per_cta * 32,
                                                       SBX_LUT_MATH_BYTES);
    img.variant = variant;
    {   // tile geometry the image was compiled for (sbx_kernel.cuh); images without the symbol use the default 8x4
        CUdeviceptr info = 0;
        size_t bytes = 0;
        int v[4] = {0, 0, 0, 0};
        if (ctx->cu->ModuleGetGlobal(&info, &bytes, img.module, "sbx_image_info") == CUDA_SUCCESS && bytes >= sizeof v &&
            ctx->cu->MemcpyDtoH(v, info, sizeof v) == CUDA_SUCCESS && v[0] > 0 && v[1] > 0 && v[2] > 0) {
            img.tile_w = v[0]; img.tile_h = v[1]; img.lanes_per_pixel = v[2]; img.hybrid_lanes = v[3] > 1 ? v[3] : 0;
        }
    }
    {
        CUdeviceptr hints = 0;
        size_t bytes = 0;
        int v[4] = {0, 0, 0, 0};
        if (ctx->cu->ModuleGetGlobal(&hints, &bytes, img.module, "sbx_image_hints") == CUDA_SUCCESS && bytes >= sizeof v &&
            ctx->cu->MemcpyDtoH(v, hints, sizeof v) == CUDA_SUCCESS) {
            if (v[0] > 0 && v[0] < 1000) img.trivial_rows_permille = v[0];
            img.uses_noise_tex = v[1] == 1;
        }
    }
    auto it = ctx->images.find(key);
    if (it != ctx->images.end() && it->second.module) ctx->cu->ModuleUnload(it->second.module);
    ctx->images[key] = img;
    return SBX_OK;
}

bool valid_shard(const sbx_shard* s, sbx_shard* out) {
    sbx_shard r = {1, 1, 0};
    if (s && s->n_parts < 0) return false;
    if (s && s->n_parts > 0) r = *s;
    if (r.stripe_rows <= 0) r.stripe_rows = 1;
    if (r.part < 0 || r.part >= r.n_parts) return false;
    *out = r;
    return true;
}

int shard_rows(const sbx_shard& s, int height) {
    int rows = 0;
    const int period = s.stripe_rows * s.n_parts;
    const int full = height / period;
    rows = full * s.stripe_rows;
    const int rem = height - full * period;            // rows of the last, partial period
    const int start = s.part * s.stripe_rows;
    if (rem > start) rows += (rem - start < s.stripe_rows) ? rem - start : s.stripe_rows;
    return rows;
}

}  // namespace

extern "C" {

const char* sbx_version(void) { return "shaderbox_b200 0.1 (sm_100a)"; }

const char* sbx_strerror(int status) {
    switch (status) {
        case SBX_OK: return "ok";
        case SBX_ERR_INVALID: return "invalid argument";
        case SBX_ERR_NO_DEVICE: return "no usable CUDA device";
        case SBX_ERR_UNKNOWN_APP: return "unknown app / kernel image missing";
        case SBX_ERR_CUDA: return "CUDA call failed";
        case SBX_ERR_COMPILE: return "app header failed to compile";
        case SBX_ERR_NOMEM: return "out of memory";
        case SBX_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

const char* sbx_last_error(sbx_ctx* ctx) { return ctx ? ctx->last_error.c_str() : g_last_error.c_str(); }

int sbx_default_params(sbx_params* p, int width, int height) {
    if (!p || width <= 0 || height <= 0) return SBX_ERR_INVALID;
    std::memset(p, 0, sizeof *p);
    p->width = width;
    p->height = height;
    p->u_time = 0.0f;
    // src/uniform_buffer.h:41-54
    p->wind_dir[0] = 0.0f; p->wind_dir[1] = 0.0f; p->wind_dir[2] = 0.2f;
    p->sun_dir[0] = 0.0f; p->sun_dir[1] = 0.0f; p->sun_dir[2] = -1.0f;
    p->sun_color[0] = 1.0f; p->sun_color[1] = 0.7f; p->sun_color[2] = 0.55f;
    p->sun_power = 8.0f;
    p->cld_march_steps = 100;
    p->illum_march_steps = 6;
    p->sigma_scattering = 0.15f;
    p->cld_coverage = 0.535f;
    p->cld_thick = 125.0f;
    p->atm_radius = 5000.0f;
    p->atm_ground_y = 4750.0f;
    // src/uniform_buffer.h:57-58
    p->fog_density = 0.1f;
    p->fog_falloff = 0.5f;
    return SBX_OK;
}

int sbx_device_count(void) {
    std::string err;
    driver_api* cu = load_driver(&err);
    int n = 0;
    if (!cu || cu->Init(0) != CUDA_SUCCESS || cu->DeviceGetCount(&n) != CUDA_SUCCESS) return 0;
    return n;
}

int sbx_create(int device, sbx_ctx** out) {
    if (!out) return SBX_ERR_INVALID;
    *out = nullptr;
    std::string err;
    driver_api* cu = load_driver(&err);
    if (!cu) { g_last_error = err; return SBX_ERR_NO_DEVICE; }
    if (cu->Init(0) != CUDA_SUCCESS) { g_last_error = "cuInit failed: no usable CUDA device"; return SBX_ERR_NO_DEVICE; }
    sbx_ctx* ctx = new sbx_ctx;
    ctx->cu = cu;
    ctx->device = device;
    if (cu->DeviceGet(&ctx->dev, device) != CUDA_SUCCESS) {
        g_last_error = "cuDeviceGet failed";
        delete ctx;
        return SBX_ERR_NO_DEVICE;
    }
    int major = 0, minor = 0;
    cu->DeviceGetAttribute(&major, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, ctx->dev);
    cu->DeviceGetAttribute(&minor, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, ctx->dev);
    cu->DeviceGetAttribute(&ctx->sm_count, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, ctx->dev);
    if (major != 10) {
        char buf[128];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; the kernel images are sm_100a only", device, major, minor);
        g_last_error = buf;
        delete ctx;
        return SBX_ERR_NO_DEVICE;
    }
    if (cu->DevicePrimaryCtxRetain(&ctx->ctx, ctx->dev) != CUDA_SUCCESS) {
        g_last_error = "cuDevicePrimaryCtxRetain failed";
        delete ctx;
        return SBX_ERR_CUDA;
    }
    {
        ctx_scope scope(ctx);
        cu->EventCreate(&ctx->ev0, CU_EVENT_DEFAULT);
        cu->EventCreate(&ctx->ev1, CU_EVENT_DEFAULT);
        cu->EventCreate(&ctx->ev2, CU_EVENT_DEFAULT);
    }
    *out = ctx;
    return SBX_OK;
}

void sbx_destroy(sbx_ctx* ctx) {
    if (!ctx) return;
    {
        ctx_scope scope(ctx);
        for (auto& kv : ctx->images)
            if (kv.second.module) ctx->cu->ModuleUnload(kv.second.module);
        if (ctx->util_module) ctx->cu->ModuleUnload(ctx->util_module);
        if (ctx->lut) ctx->cu->MemFree(ctx->lut);
        if (ctx->hash_tab) ctx->cu->MemFree(ctx->hash_tab);
        if (ctx->times_ring) ctx->cu->MemFreeHost(ctx->times_ring);
        for (CUevent e : ctx->times_ev) if (e) ctx->cu->EventDestroy(e);
        if (ctx->frame) ctx->cu->MemFree(ctx->frame);
        if (ctx->done_counters) ctx->cu->MemFree(ctx->done_counters);
        for (CUdeviceptr v : ctx->noise_vol) if (v) ctx->cu->MemFree(v);
        if (ctx->ev0) ctx->cu->EventDestroy(ctx->ev0);
        if (ctx->ev1) ctx->cu->EventDestroy(ctx->ev1);
        if (ctx->ev2) ctx->cu->EventDestroy(ctx->ev2);
    }
    ctx->cu->DevicePrimaryCtxRelease(ctx->dev);
    delete ctx;
}

int sbx_compile_app(sbx_ctx* ctx, const char* app_header_path, const char* app_name, const char* image_out_path) {
    if (!app_header_path || !app_name) return SBX_ERR_INVALID;
    std::string cubin, log;
    // extra -D options for the image, e.g. SBX_COMPILE_DEFINES="SBX_MIN_CTAS_PER_SM=4;SBX_WARPS_PER_CTA=4" (launch shape
    // of the pixel-loop kernel, sbx_kernel.cuh); the per-app defaults used by the build are in csrc/Makefile
    std::vector<std::string> defines;
    if (const char* env = std::getenv("SBX_COMPILE_DEFINES")) {
        std::stringstream ss(env);
        for (std::string d; std::getline(ss, d, ';');)
            if (!d.empty()) defines.push_back(d);
    }
    const int st = sbx::compile_app_header(app_header_path, app_name, defines, &cubin, &log);
    if (st != SBX_OK) {
        if (ctx) ctx->last_error = log; else g_last_error = log;
        return st;
    }
    if (image_out_path) {
        std::ofstream f(image_out_path, std::ios::binary);
        if (!f) { g_last_error = std::string("cannot write ") + image_out_path; return SBX_ERR_INVALID; }
        f.write(cubin.data(), (std::streamsize)cubin.size());
    }
    if (ctx) {
        ctx_scope scope(ctx);
        const std::string key = std::string(app_name) + "/plugin";
        const int bs = bind_image(ctx, key, cubin, "plugin");
        if (bs != SBX_OK) return bs;
        ctx->last_error = log;   // warnings, if any
    } else {
        g_last_error = log;
    }
    return SBX_OK;
}

int sbx_load_app(sbx_ctx* ctx, const char* app_name, const char* variant) {
    if (!ctx || !app_name) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    const char* order_default[] = {"native", "plugin"};
    std::vector<std::string> order;
    if (variant && *variant) order.push_back(variant);
    else for (const char* v : order_default) order.push_back(v);
    for (const auto& v : order) {
        const std::string key = std::string(app_name) + "/" + v;
        auto it = ctx->images.find(key);
        if (it == ctx->images.end()) {
            std::string bin;
            const std::string path = sbx::library_dir() + "/images/" + app_name + "." + v + ".cubin";
            if (!read_file(path, &bin)) continue;
            const int st = bind_image(ctx, key, bin, v);
            if (st != SBX_OK) return st;
            it = ctx->images.find(key);
        }
        ctx->current = &it->second;
        ctx->current_coop = nullptr;
        ctx->current_coop2 = nullptr;
        ctx->current_hybrid = nullptr;
        ctx->current_app = app_name;
        if (!(variant && *variant) && v == "native") {
            // default selection only: the cooperative builds of the same scene kernel, if shipped
            for (const char* cv : {"coop", "coop2", "hybrid"}) {
                const std::string ckey = std::string(app_name) + "/" + cv;
                auto ct = ctx->images.find(ckey);
                if (ct == ctx->images.end()) {
                    std::string cbin;
                    if (read_file(sbx::library_dir() + "/images/" + app_name + "." + cv + ".cubin", &cbin) &&
                        bind_image(ctx, ckey, cbin, cv) == SBX_OK)
                        ct = ctx->images.find(ckey);
                }
                if (ct != ctx->images.end())
                    (!std::strcmp(cv, "coop") ? ctx->current_coop : !std::strcmp(cv, "coop2") ? ctx->current_coop2 : ctx->current_hybrid) = &ct->second;
            }
        }
        return SBX_OK;
    }
    return ctx->fail(SBX_ERR_UNKNOWN_APP, "no kernel image for %s (variant %s) under %s/images", app_name,
                     variant && *variant ? variant : "native|plugin", sbx::library_dir().c_str());
}

int sbx_shard_rows(const sbx_shard* shard, int height) {
    sbx_shard s;
    if (height <= 0 || !valid_shard(shard, &s)) return SBX_ERR_INVALID;
    return shard_rows(s, height);
}

int sbx_set_option(sbx_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return SBX_ERR_INVALID;
    if (!std::strcmp(key, "hash_table_log2")) {
        if (value < 9 || value > 22) return SBX_ERR_INVALID;
        ctx->opt_hash_log2 = value;
        return SBX_OK;
    }
    if (!std::strcmp(key, "use_hash_table")) { ctx->opt_use_hash = value ? 1 : 0; return SBX_OK; }
    if (!std::strcmp(key, "host_zero_copy")) { ctx->opt_zero_copy = value ? 1 : 0; return SBX_OK; }
    if (!std::strcmp(key, "coop_waves_x100")) { if (value < 0) return SBX_ERR_INVALID; ctx->opt_coop_waves_x100 = value; return SBX_OK; }
    if (!std::strcmp(key, "trivial_rows_last")) { ctx->opt_trivial_rows_last = value ? 1 : 0; return SBX_OK; }
    if (!std::strcmp(key, "record_events")) { ctx->opt_record_events = value ? 1 : 0; return SBX_OK; }
    if (!std::strcmp(key, "tail_waves_x100")) { if (value < 0) return SBX_ERR_INVALID; ctx->opt_tail_waves_x100 = value; return SBX_OK; }
    if (!std::strcmp(key, "tail_max_waves_x100")) { if (value < 0) return SBX_ERR_INVALID; ctx->opt_tail_max_waves_x100 = value; return SBX_OK; }
    return SBX_ERR_INVALID;
}

// what one render launch covers and where it goes (the public entry points fill one of these)
struct launch_job {
    const sbx_shard* rows = nullptr;   // row stripes (NULL = every row)
    int col_parts = 1, col_part = 0;   // tile-column interleave (frame outputs only)
    int out_is_frame = 0, out_rgba8 = 0;
    const float* dev_times = nullptr;  // sequence launch
    int n_frames = 1;
    unsigned* done_flag = nullptr;     // completion signal
    unsigned done_value = 0;
};

static void plan_region(sbx_region* r, int width, int tile_w, int tile_h, int col_parts, int row0, int rows, int warps_per_cta,
                        int first_row = 0) {
    const int tiles_x = (width + tile_w - 1) / tile_w;
    r->first_tile_row = rows > 0 ? (first_row / tile_h) % ((rows + tile_h - 1) / tile_h) : 0;
    r->tiles_per_row = col_parts > 1 ? (tiles_x + col_parts - 1) / col_parts : tiles_x;
    r->row0 = row0;
    r->rows = rows;
    r->tile_rows = (rows + tile_h - 1) / tile_h;
    r->warps = r->tile_rows * r->tiles_per_row;
    // the grid is rounded up to whole CTAs: the largest warp index is < warps + warps_per_cta
    const unsigned long long nmax = (unsigned long long)r->warps + (unsigned long long)warps_per_cta;
    r->magic = (r->tiles_per_row > 0 && nmax * (unsigned long long)r->tiles_per_row < (1ull << 40))
                   ? ((1ull << 40) + (unsigned long long)r->tiles_per_row - 1) / (unsigned long long)r->tiles_per_row : 0ull;
}

static int render_launch(sbx_ctx* ctx, const sbx_params* p, const launch_job& job, float* dev_rgba, void* stream_) {
    if (!ctx || !p || !dev_rgba || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    if (!ctx->current) return ctx->fail(SBX_ERR_UNKNOWN_APP, "sbx_load_app was not called");
    sbx_shard s;
    if (!valid_shard(job.rows, &s)) return ctx->fail(SBX_ERR_INVALID, "bad shard");
    if (job.col_parts < 1 || job.col_part < 0 || job.col_part >= job.col_parts) return ctx->fail(SBX_ERR_INVALID, "bad tile shard");
    if (job.col_parts > 1 && !job.out_is_frame) return ctx->fail(SBX_ERR_INVALID, "tile shards render into a full frame only");
    const int n_frames = job.n_frames;
    CUstream stream = (CUstream)stream_;
    ctx_scope scope(ctx);
    ctx->timing.launches = 0;
    int st = ensure_tables(ctx, stream, ctx->current->variant != "plugin");
    if (st != SBX_OK) return st;

    sbx_launch L;
    std::memset(&L, 0, sizeof L);
    L.p = *p;
    L.stripe_rows = s.stripe_rows;
    L.n_parts = s.n_parts;
    L.part = s.part;
    L.local_rows = shard_rows(s, p->height);
    L.col_parts = job.col_parts;
    L.col_part = job.col_part;
    kernel_image* img = ctx->current;
    int rows_tail = 0;   // rows of the launch marched by the hybrid image's second region
    if (L.local_rows > 0 && (ctx->current_coop || ctx->current_coop2 || ctx->current_hybrid)) {
        // A frame (or one rank's share of it) that is only a few waves of resident warps ends in a long tail of single
        // warps still marching.  The hybrid image marches the LAST rows of the launch (about one wave of warps) with 4
        // lanes per pixel -- pieces 4 times shorter, which drain 4 times faster -- and everything before them with one
        // lane per pixel, so only the tail pays the cooperative march's overhead (multi-GPU).  Without a
        // hybrid image the whole launch switches to a cooperative image when it is small: measured on one rank's share
        // of CLOUDS 1080p (tools/shard_time.py): 2.3 waves -> P=4 wins, 4.6 -> P=2, 9 -> P=1.
        const int cols = job.col_parts > 1 ? job.col_parts : 1;
        const long long tiles_x = ((p->width + img->tile_w - 1) / img->tile_w + cols - 1) / cols;
        const long long warps = tiles_x * ((L.local_rows + img->tile_h - 1) / img->tile_h) * n_frames;
        const long long resident = (long long)ctx->sm_count * img->blocks_per_sm * img->warps_per_cta;
        if (ctx->current_hybrid && n_frames == 1 && ctx->opt_tail_waves_x100 > 0) {
            if (warps * 100 < resident * ctx->opt_tail_max_waves_x100) {
                // tail = opt_tail_waves waves of one-lane warps' worth of rows, in whole 8x4 tile rows
                long long tail_warps = resident * ctx->opt_tail_waves_x100 / 100;
                long long tile_rows = (tail_warps + tiles_x - 1) / tiles_x;
                rows_tail = (int)std::min<long long>((long long)L.local_rows, tile_rows * SBX_TILE_H);
                if (L.local_rows - rows_tail < SBX_TILE_H) rows_tail = L.local_rows;
                rows_tail = L.local_rows - ((L.local_rows - rows_tail) / SBX_TILE_H) * SBX_TILE_H;   // region 0 is whole tile rows
                img = ctx->current_hybrid;
            }
        } else if (ctx->current_coop && warps * 100 < resident * ctx->opt_coop_waves_x100) img = ctx->current_coop;
        else if (ctx->current_coop2 && warps * 100 < resident * ctx->opt_coop_waves_x100 * 2) img = ctx->current_coop2;
    }
    // the tile checkerboard is defined on 8-pixel columns: every image with 8-wide tiles (one lane per pixel, 4 lanes,
    // hybrid) cuts the same parts; the 2-lane image (16-wide tiles) would not
    if (job.col_parts > 1 && img->tile_w != SBX_TILE_W) { img = ctx->current; rows_tail = 0; }
    if (L.local_rows > 0) {
        // rows the image declares trivial (a share of the frame's bottom) go last: the launch starts at this shard's first
        // row at or above that frame row -- counted in FRAME rows, so every part of a striped frame starts on the same line
        const int first_row = ctx->opt_trivial_rows_last && rows_tail == 0 && img->trivial_rows_permille > 0
                                  ? shard_rows(s, (int)((long long)p->height * img->trivial_rows_permille / 1000)) : 0;
        plan_region(&L.reg[0], p->width, img->tile_w, img->tile_h, job.col_parts, 0, L.local_rows - rows_tail, img->warps_per_cta, first_row);
        if (rows_tail > 0)
            plan_region(&L.reg[1], p->width, 32 / img->hybrid_lanes, 1, job.col_parts, L.local_rows - rows_tail, rows_tail, img->warps_per_cta);
    }
    L.out = dev_rgba;
    L.out_is_frame = job.out_is_frame;
    L.out_rgba8 = job.out_rgba8;
    L.times = job.dev_times;
    L.hash_tab = (const float4*)ctx->hash_tab;
    L.hash_bias = SBX_HASH_MAGIC_BITS + ctx->hash_lo;
    L.hash_len = ctx->hash_len;
    L.hash_span = ctx->hash_len;
    L.lut = (const void*)ctx->lut;
    L.trace = (unsigned long long*)(uintptr_t)ctx->trace;
    if (job.done_flag) {
        if (!ctx->done_counters) {
            SBX_TRY(ctx->cu->MemAlloc(&ctx->done_counters, 64 * sizeof(unsigned)), "cuMemAlloc(done counters)");
            SBX_TRY(ctx->cu->MemsetD32(ctx->done_counters, 0u, 64), "cuMemsetD32(done counters)");
        }
        L.done_counter = (unsigned*)(uintptr_t)(ctx->done_counters + (ctx->done_seq++ % 64u) * sizeof(unsigned));
        L.done_flag = job.done_flag;
        L.done_value = job.done_value;
    }

    const long long warps = (long long)L.reg[0].warps + L.reg[1].warps;
    // a launch without pixels still publishes its completion flag (one idle CTA)
    if (warps == 0 && !job.done_flag) return SBX_OK;
    const unsigned grid = (unsigned)std::max<long long>(1, (warps + img->warps_per_cta - 1) / img->warps_per_cta);
    if (img->uses_noise_tex && !ctx->noise_vol[0])
        return ctx->fail(SBX_ERR_INVALID, "%s samples the 3-D noise textures: call sbx_set_noise_volumes first", ctx->current_app.c_str());
    void* args[] = {&L, &ctx->tex};   // (the second parameter exists only in images that declare it)
    // the two timing events cost ~2 us of stream time each: a caller that times its own launches switches them off
    const bool timed = ctx->opt_record_events || !job.out_is_frame;
    if (timed) SBX_TRY(ctx->cu->EventRecord(ctx->ev0, stream), "cuEventRecord");
    SBX_TRY(ctx->cu->LaunchKernel(img->render, grid, (unsigned)n_frames, 1, (unsigned)img->warps_per_cta * 32, 1, 1,
                                  SBX_LUT_MATH_BYTES, stream, args, nullptr),
            "launch sbx_render");
    if (timed) SBX_TRY(ctx->cu->EventRecord(ctx->ev1, stream), "cuEventRecord");
    ctx->timing.launches += 1;
    ctx->timing.grid_blocks = (int)grid;
    ctx->timing.block_threads = img->warps_per_cta * 32;
    ctx->timing.regs_per_thread = img->regs;
    ctx->timing.blocks_per_sm = img->blocks_per_sm;
    ctx->timing.lanes_per_pixel = img->lanes_per_pixel;
    ctx->timing.tail_rows = rows_tail;
    ctx->timing.tail_lanes_per_pixel = rows_tail > 0 ? img->hybrid_lanes : 0;
    ctx->timing.kernel_ms = timed ? -1.0f : 0.0f;   // -1: resolved lazily by sbx_last_timing
    ctx->timing.d2h_ms = 0.0f;
    return SBX_OK;
}

// Device-visible alias of a host frame, if the caller's buffer is pinned and mapped (cudaHostAlloc /
// cuMemHostAlloc / a pinned torch tensor): 0 otherwise.
static CUdeviceptr mapped_host_alias(sbx_ctx* ctx, const void* host) {
    unsigned mem_type = 0;
    if (ctx->cu->PointerGetAttribute(&mem_type, CU_POINTER_ATTRIBUTE_MEMORY_TYPE, (CUdeviceptr)(uintptr_t)host) != CUDA_SUCCESS)
        return 0;
    if (mem_type != CU_MEMORYTYPE_HOST) return 0;
    CUdeviceptr dptr = 0;
    if (ctx->cu->PointerGetAttribute(&dptr, CU_POINTER_ATTRIBUTE_DEVICE_POINTER, (CUdeviceptr)(uintptr_t)host) != CUDA_SUCCESS)
        return 0;
    return dptr;
}

int sbx_render_device(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, float* dev_rgba, void* stream) {
    launch_job job;
    job.rows = shard;
    return render_launch(ctx, p, job, dev_rgba, stream);
}

int sbx_render_sequence_device(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, const float* times, int n_frames,
                               float* dev_rgba, void* stream_) {
    if (!ctx || !times || n_frames < 1 || n_frames > 65535) return SBX_ERR_INVALID;
    CUstream stream = (CUstream)stream_;
    CUdeviceptr dev_times = 0;
    const size_t bytes = (size_t)n_frames * sizeof(float);
    {
        ctx_scope scope(ctx);
        if (!ctx->times_ring) {
            void* ring = nullptr;
            if (ctx->cu->MemHostAlloc(&ring, (size_t)sbx_ctx::kTimeSlots * sbx_ctx::kTimeSlotFloats * sizeof(float), 0) != CUDA_SUCCESS)
                return ctx->fail(SBX_ERR_NOMEM, "cuMemHostAlloc(times ring) failed");
            ctx->times_ring = (float*)ring;
            for (CUevent& e : ctx->times_ev) SBX_TRY(ctx->cu->EventCreate(&e, CU_EVENT_DISABLE_TIMING), "cuEventCreate");
        }
        // the slot is free once the copy that last used it has run (8 sequence launches ago: normally long done)
        const unsigned slot = ctx->times_seq++ % (unsigned)sbx_ctx::kTimeSlots;
        if (ctx->times_seq > (unsigned)sbx_ctx::kTimeSlots) SBX_TRY(ctx->cu->EventSynchronize(ctx->times_ev[slot]), "cuEventSynchronize(times slot)");
        float* staged = ctx->times_ring + (size_t)slot * sbx_ctx::kTimeSlotFloats;
        std::memcpy(staged, times, bytes);          // the caller's array is consumed here
        SBX_TRY(ctx->cu->MemAllocAsync(&dev_times, bytes, stream), "cuMemAllocAsync(times)");
        SBX_TRY(ctx->cu->MemcpyHtoDAsync(dev_times, staged, bytes, stream), "cuMemcpyHtoDAsync(times)");
        SBX_TRY(ctx->cu->EventRecord(ctx->times_ev[slot], stream), "cuEventRecord(times slot)");
    }
    launch_job job;
    job.rows = shard;
    job.dev_times = (const float*)(uintptr_t)dev_times;
    job.n_frames = n_frames;
    const int st = render_launch(ctx, p, job, dev_rgba, stream_);
    ctx_scope scope(ctx);
    ctx->cu->MemFreeAsync(dev_times, stream);       // stream-ordered: after the kernel that reads it
    return st;
}

int sbx_render_sequence_host(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, const float* times, int n_frames,
                             float* host_rgba) {
    if (!ctx || !p || !host_rgba || !times || n_frames < 1 || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    sbx_shard s;
    if (!valid_shard(shard, &s)) return ctx->fail(SBX_ERR_INVALID, "bad shard");
    const size_t bytes = (size_t)n_frames * (size_t)shard_rows(s, p->height) * (size_t)p->width * 4 * sizeof(float);
    if (bytes == 0) return SBX_OK;
    CUdeviceptr alias = 0;
    {
        ctx_scope scope(ctx);
        if (ctx->opt_zero_copy && ((uintptr_t)host_rgba & 15u) == 0) alias = mapped_host_alias(ctx, host_rgba);
        if (!alias && bytes > ctx->frame_bytes) {
            if (ctx->frame) ctx->cu->MemFree(ctx->frame);
            ctx->frame = 0;
            ctx->frame_bytes = 0;
            if (ctx->cu->MemAlloc(&ctx->frame, bytes) != CUDA_SUCCESS) return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc(%zu) failed", bytes);
            ctx->frame_bytes = bytes;
        }
    }
    int st = sbx_render_sequence_device(ctx, p, &s, times, n_frames, (float*)(alias ? alias : ctx->frame), nullptr);
    if (st != SBX_OK) return st;
    ctx_scope scope(ctx);
    if (!alias) SBX_TRY(ctx->cu->MemcpyDtoHAsync(host_rgba, ctx->frame, bytes, nullptr), "cuMemcpyDtoHAsync");
    SBX_TRY(ctx->cu->EventRecord(ctx->ev2, nullptr), "cuEventRecord");
    SBX_TRY(ctx->cu->StreamSynchronize(nullptr), "cuStreamSynchronize");
    float ms = 0.0f;
    if (ctx->cu->EventElapsedTime(&ms, ctx->ev1, ctx->ev2) == CUDA_SUCCESS) ctx->timing.d2h_ms = ms;
    ctx->timing.zero_copy = alias ? 1 : 0;
    return SBX_OK;
}

int sbx_render_frame(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, float* dev_frame, void* stream) {
    launch_job job;
    job.rows = shard;
    job.out_is_frame = 1;
    return render_launch(ctx, p, job, dev_frame, stream);
}

int sbx_render_frame_part(sbx_ctx* ctx, const sbx_params* p, const sbx_frame_part* part, float* dev_frame, void* stream) {
    if (!part) return SBX_ERR_INVALID;
    launch_job job;
    job.out_is_frame = 1;
    sbx_shard rows = part->rows;
    job.rows = &rows;
    job.col_parts = part->tile_parts > 1 ? part->tile_parts : 1;
    job.col_part = part->tile_parts > 1 ? part->tile_part : 0;
    if (part->tile_parts > 1 && (part->tile_part < 0 || part->tile_part >= part->tile_parts))
        return ctx ? ctx->fail(SBX_ERR_INVALID, "bad tile part") : SBX_ERR_INVALID;
    job.done_flag = part->done_flag;
    job.done_value = part->done_value;
    return render_launch(ctx, p, job, dev_frame, stream);
}

int sbx_stream_wait_flags(sbx_ctx* ctx, const unsigned* dev_flags, int n, unsigned value, void* stream) {
    if (!ctx || !dev_flags || n < 0) return SBX_ERR_INVALID;
    if (n == 0) return SBX_OK;
    if (n > 256) return ctx->fail(SBX_ERR_INVALID, "at most 256 flags per wait");
    ctx_scope scope(ctx);
    // one submission for all n waits (cuStreamBatchMemOp)
    CUstreamBatchMemOpParams ops[256];
    std::memset(ops, 0, sizeof(ops[0]) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        ops[i].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
        ops[i].waitValue.address = (CUdeviceptr)(uintptr_t)(dev_flags + i);
        ops[i].waitValue.value = value;
        ops[i].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
    }
    SBX_TRY(ctx->cu->StreamBatchMemOp((CUstream)stream, (unsigned)n, ops, 0), "cuStreamBatchMemOp(wait flags)");
    return SBX_OK;
}

int sbx_set_noise_volumes(sbx_ctx* ctx, const float* host_rgba_a, const float* host_rgba_b, int size) {
    if (!ctx || !host_rgba_a || !host_rgba_b || size < 2 || size > 1024) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    const int padded = size + 2;
    const int pitch_x = (padded + 3) / 4 * 4;                   // rows of whole 16-byte units (TMA strides)
    const size_t floats = (size_t)pitch_x * padded * padded;
    std::vector<float> staging(floats);
    SBX_TRY(ctx->cu->CtxSynchronize(), "cuCtxSynchronize");   // launches still sampling the old volumes, on any stream
    const float* src[2] = {host_rgba_a, host_rgba_b};
    for (int t = 0; t < 2; ++t) {
        // the .r channel with a one-texel apron of wrapped neighbours: padded[k] = texel[(k - 1) mod size]
        std::fill(staging.begin(), staging.end(), 0.0f);
        for (int z = 0; z < padded; ++z)
            for (int y = 0; y < padded; ++y) {
                const size_t sz = (size_t)((z - 1 + size) % size), sy = (size_t)((y - 1 + size) % size);
                float* row = staging.data() + ((size_t)z * padded + y) * pitch_x;
                for (int x = 0; x < padded; ++x) row[x] = src[t][((sz * size + sy) * size + (size_t)((x - 1 + size) % size)) * 4];
            }
        if (ctx->noise_vol[t]) { ctx->cu->MemFree(ctx->noise_vol[t]); ctx->noise_vol[t] = 0; }
        if (ctx->cu->MemAlloc(&ctx->noise_vol[t], floats * sizeof(float)) != CUDA_SUCCESS) return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc(noise volume) failed");
        SBX_TRY(ctx->cu->MemcpyHtoD(ctx->noise_vol[t], staging.data(), floats * sizeof(float)), "cuMemcpyHtoD(noise volume)");
        // TMA descriptor: rank 3 fp32, x fastest, 8 x 4 x 4 boxes (x starts on 16-byte boundaries), no swizzle, zero fill outside
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)padded, (cuuint64_t)padded, (cuuint64_t)padded};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch_x * sizeof(float), (cuuint64_t)pitch_x * padded * sizeof(float)};
        const cuuint32_t box[3] = {8, 4, 4}, elem[3] = {1, 1, 1};
        SBX_TRY(ctx->cu->TensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)(uintptr_t)ctx->noise_vol[t], dims, strides, box, elem,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
                "cuTensorMapEncodeTiled");
        static_assert(sizeof(CUtensorMap) == sizeof(ctx->tex.map[0]), "CUtensorMap is 128 bytes");
        std::memcpy(ctx->tex.map[t], &map, sizeof map);     // travels as a __grid_constant__ kernel parameter
        ctx->tex.vol[t] = (const float*)(uintptr_t)ctx->noise_vol[t];
    }
    ctx->tex.size = size;
    ctx->tex.pitch_x = pitch_x;
    ctx->tex.pitch_xy = pitch_x * padded;
    return SBX_OK;
}

int sbx_stream_write_flag(sbx_ctx* ctx, unsigned* dev_flag, unsigned value, void* stream) {
    if (!ctx || !dev_flag) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    // default flags: the write is preceded by a system-scope fence over the stream's earlier work
    SBX_TRY(ctx->cu->StreamWriteValue32((CUstream)stream, (CUdeviceptr)(uintptr_t)dev_flag, value, CU_STREAM_WRITE_VALUE_DEFAULT),
            "cuStreamWriteValue32");
    return SBX_OK;
}

int sbx_host_alloc(sbx_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out || bytes == 0) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    void* h = nullptr;
    if (ctx->cu->MemHostAlloc(&h, bytes, CU_MEMHOSTALLOC_PORTABLE | CU_MEMHOSTALLOC_DEVICEMAP) != CUDA_SUCCESS)
        return ctx->fail(SBX_ERR_NOMEM, "cuMemHostAlloc(%zu) failed", bytes);
    *out = h;
    return SBX_OK;
}

int sbx_host_free(sbx_ctx* ctx, void* host) {
    if (!ctx || !host) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemFreeHost(host), "cuMemFreeHost");
    return SBX_OK;
}

int sbx_frame_alloc(sbx_ctx* ctx, size_t bytes, float** out) {
    if (!ctx || !out || bytes == 0) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    CUdeviceptr d = 0;
    if (ctx->cu->MemAlloc(&d, bytes) != CUDA_SUCCESS) return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc(%zu) failed", bytes);
    if ((bytes & 3u) == 0) ctx->cu->MemsetD32(d, 0u, bytes / 4);   // frames (and completion flags kept behind them) start zeroed
    *out = (float*)(uintptr_t)d;
    return SBX_OK;
}

int sbx_frame_free(sbx_ctx* ctx, float* dev) {
    if (!ctx || !dev) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemFree((CUdeviceptr)(uintptr_t)dev), "cuMemFree");
    return SBX_OK;
}

int sbx_frame_export(sbx_ctx* ctx, const float* dev, unsigned char handle[SBX_IPC_HANDLE_BYTES]) {
    if (!ctx || !dev || !handle) return SBX_ERR_INVALID;
    static_assert(sizeof(CUipcMemHandle) == SBX_IPC_HANDLE_BYTES, "CUipcMemHandle is 64 bytes");
    ctx_scope scope(ctx);
    CUipcMemHandle h;
    SBX_TRY(ctx->cu->IpcGetMemHandle(&h, (CUdeviceptr)(uintptr_t)dev), "cuIpcGetMemHandle");
    std::memcpy(handle, &h, sizeof h);
    return SBX_OK;
}

int sbx_frame_import(sbx_ctx* ctx, const unsigned char handle[SBX_IPC_HANDLE_BYTES], float** out) {
    if (!ctx || !handle || !out) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    CUipcMemHandle h;
    std::memcpy(&h, handle, sizeof h);
    CUdeviceptr d = 0;
    SBX_TRY(ctx->cu->IpcOpenMemHandle(&d, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS), "cuIpcOpenMemHandle");
    *out = (float*)(uintptr_t)d;
    return SBX_OK;
}

int sbx_frame_release(sbx_ctx* ctx, float* imported) {
    if (!ctx || !imported) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->IpcCloseMemHandle((CUdeviceptr)(uintptr_t)imported), "cuIpcCloseMemHandle");
    return SBX_OK;
}

int sbx_host_frame_register(sbx_ctx* ctx, void* host, size_t bytes, float** dev_alias_out) {
    if (!ctx || !host || !dev_alias_out || bytes == 0) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemHostRegister(host, bytes, CU_MEMHOSTREGISTER_PORTABLE | CU_MEMHOSTREGISTER_DEVICEMAP), "cuMemHostRegister");
    CUdeviceptr d = 0;
    const CUresult r = ctx->cu->MemHostGetDevicePointer(&d, host, 0);
    if (r != CUDA_SUCCESS) {
        ctx->cu->MemHostUnregister(host);
        return ctx->check(r, "cuMemHostGetDevicePointer");
    }
    *dev_alias_out = (float*)(uintptr_t)d;
    return SBX_OK;
}

int sbx_host_frame_unregister(sbx_ctx* ctx, void* host) {
    if (!ctx || !host) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemHostUnregister(host), "cuMemHostUnregister");
    return SBX_OK;
}

int sbx_frame_read(sbx_ctx* ctx, const float* dev, float* host, size_t bytes, void* stream) {
    if (!ctx || !dev || !host) return SBX_ERR_INVALID;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemcpyDtoHAsync(host, (CUdeviceptr)(uintptr_t)dev, bytes, (CUstream)stream), "cuMemcpyDtoHAsync");
    SBX_TRY(ctx->cu->StreamSynchronize((CUstream)stream), "cuStreamSynchronize");
    return SBX_OK;
}

// one frame to host memory, float4 or packed 8-bit pixels
static int render_host(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, void* host, int rgba8) {
    if (!ctx || !p || !host || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    sbx_shard s;
    if (!valid_shard(shard, &s)) return ctx->fail(SBX_ERR_INVALID, "bad shard");
    const size_t bytes = (size_t)shard_rows(s, p->height) * (size_t)p->width * (rgba8 ? 4 : 4 * sizeof(float));
    if (bytes == 0) return SBX_OK;
    CUdeviceptr alias = 0;
    {
        ctx_scope scope(ctx);
        // Pinned + mapped destination: the kernel's pixel stores go straight over PCIe into the
        // caller's frame while other pixels are still marching (the frame is write-only and each warp
        // writes whole lines / sectors), so there is no separate device->host copy to wait for.
        if (ctx->opt_zero_copy && ((uintptr_t)host & 15u) == 0) alias = mapped_host_alias(ctx, host);
        if (!alias && bytes > ctx->frame_bytes) {
            if (ctx->frame) ctx->cu->MemFree(ctx->frame);
            ctx->frame = 0;
            ctx->frame_bytes = 0;
            if (ctx->cu->MemAlloc(&ctx->frame, bytes) != CUDA_SUCCESS) return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc(%zu) failed", bytes);
            ctx->frame_bytes = bytes;
        }
    }
    launch_job job;
    job.rows = &s;
    job.out_rgba8 = rgba8;
    int st = render_launch(ctx, p, job, (float*)(alias ? alias : ctx->frame), nullptr);
    if (st != SBX_OK) return st;
    ctx_scope scope(ctx);
    if (!alias) SBX_TRY(ctx->cu->MemcpyDtoHAsync(host, ctx->frame, bytes, nullptr), "cuMemcpyDtoHAsync");
    SBX_TRY(ctx->cu->EventRecord(ctx->ev2, nullptr), "cuEventRecord");
    SBX_TRY(ctx->cu->StreamSynchronize(nullptr), "cuStreamSynchronize");
    float ms = 0.0f;
    if (ctx->cu->EventElapsedTime(&ms, ctx->ev1, ctx->ev2) == CUDA_SUCCESS) ctx->timing.d2h_ms = ms;
    ctx->timing.zero_copy = alias ? 1 : 0;
    return SBX_OK;
}

int sbx_render_host(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, float* host_rgba) {
    return render_host(ctx, p, shard, host_rgba, 0);
}

int sbx_render_host_rgba8(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, unsigned char* host_rgba8) {
    return render_host(ctx, p, shard, host_rgba8, 1);
}

int sbx_render_device_rgba8(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard, unsigned char* dev_rgba8, void* stream) {
    if (((uintptr_t)dev_rgba8 & 3u) != 0) return ctx ? ctx->fail(SBX_ERR_INVALID, "rgba8 frame must be 4-byte aligned") : SBX_ERR_INVALID;
    launch_job job;
    job.rows = shard;
    job.out_rgba8 = 1;
    return render_launch(ctx, p, job, (float*)dev_rgba8, stream);
}

int sbx_unshard_device(sbx_ctx* ctx, int width, int height, const sbx_shard* shard, const float* dev_part,
                       float* dev_frame, void* stream_) {
    if (!ctx || !dev_part || !dev_frame || width <= 0 || height <= 0) return SBX_ERR_INVALID;
    sbx_shard s;
    if (!valid_shard(shard, &s)) return ctx->fail(SBX_ERR_INVALID, "bad shard");
    ctx_scope scope(ctx);
    int st = load_util(ctx);
    if (st != SBX_OK) return st;
    int rows = shard_rows(s, height);
    if (rows == 0) return SBX_OK;
    const long long total = (long long)rows * width;
    void* args[] = {(void*)&dev_part, (void*)&dev_frame, &width, &rows, &s.stripe_rows, &s.n_parts, &s.part};
    SBX_TRY(ctx->cu->LaunchKernel(ctx->k_unshard, (unsigned)((total + 255) / 256), 1, 1, 256, 1, 1, 0,
                                  (CUstream)stream_, args, nullptr),
            "launch sbx_unshard_kernel");
    return SBX_OK;
}

int sbx_bake_noise_volume_device(sbx_ctx* ctx, int size, int z0, int nz, float* dev_rgba, void* stream_) {
    if (!ctx || !dev_rgba || size <= 0 || size > 2048 || z0 < 0 || nz < 0 || z0 + nz > size) return SBX_ERR_INVALID;
    if (nz == 0) return SBX_OK;
    CUstream stream = (CUstream)stream_;
    ctx_scope scope(ctx);
    int st = ensure_tables(ctx, stream);
    if (st != SBX_OK) return st;
    sbx_launch L;
    std::memset(&L, 0, sizeof L);
    L.lut = (const void*)ctx->lut;
    const long long total = (long long)size * size * nz;
    void* args[] = {&L, &dev_rgba, &size, &z0, &nz};
    SBX_TRY(ctx->cu->EventRecord(ctx->ev0, stream), "cuEventRecord");
    SBX_TRY(ctx->cu->LaunchKernel(ctx->k_bake, (unsigned)((total + 127) / 128), 1, 1, 128, 1, 1, SBX_LUT_MATH_BYTES, stream, args, nullptr),
            "launch sbx_bake_volume_kernel");
    SBX_TRY(ctx->cu->EventRecord(ctx->ev1, stream), "cuEventRecord");
    ctx->timing.launches = 1;
    ctx->timing.grid_blocks = (int)((total + 127) / 128);
    ctx->timing.block_threads = 128;
    ctx->timing.kernel_ms = -1.0f;
    return SBX_OK;
}

int sbx_bake_noise_volume_host(sbx_ctx* ctx, int size, int z0, int nz, float* host_rgba) {
    if (!ctx || !host_rgba || size <= 0 || nz < 0) return SBX_ERR_INVALID;
    const size_t bytes = (size_t)size * size * (size_t)nz * 4 * sizeof(float);
    if (bytes == 0) return SBX_OK;
    {
        ctx_scope scope(ctx);
        if (bytes > ctx->frame_bytes) {
            if (ctx->frame) ctx->cu->MemFree(ctx->frame);
            ctx->frame = 0;
            ctx->frame_bytes = 0;
            if (ctx->cu->MemAlloc(&ctx->frame, bytes) != CUDA_SUCCESS) return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc(%zu) failed", bytes);
            ctx->frame_bytes = bytes;
        }
    }
    int st = sbx_bake_noise_volume_device(ctx, size, z0, nz, (float*)(uintptr_t)ctx->frame, nullptr);
    if (st != SBX_OK) return st;
    ctx_scope scope(ctx);
    SBX_TRY(ctx->cu->MemcpyDtoHAsync(host_rgba, ctx->frame, bytes, nullptr), "cuMemcpyDtoHAsync");
    SBX_TRY(ctx->cu->StreamSynchronize(nullptr), "cuStreamSynchronize");
    return SBX_OK;
}

// util/ddsvolgen/src/ddsvolgen.cpp:13-18,69-92 -- struct DDS { DWORD magic; DDS_HEADER; DDS_HEADER_DXT10 } with the
// constants of DirectXTex's DDS.h (lib/DirectXTex/DirectXTex/DDS.h:40,156-192; published DDS file format) and the
// field values ddsvolgen sets, quirks included (dwCaps carries the "cubemap" bit, pitch is (size*16 + 7) / 8).
int sbx_dds_volume_header(int size, unsigned char* out, int capacity) {
    if (!out || size <= 0 || capacity < 148) return SBX_ERR_INVALID;
    uint32_t w[37];
    std::memset(w, 0, sizeof w);
    w[0] = 0x20534444u;                                  // DDS_MAGIC "DDS "
    w[1] = 124u;                                         // header.dwSize = sizeof(DDS_HEADER)
    w[2] = 0x00001007u | 0x00800000u | 0x00000008u;      // DDS_HEADER_FLAGS_TEXTURE | _VOLUME | _PITCH
    w[3] = (uint32_t)size;                               // dwHeight
    w[4] = (uint32_t)size;                               // dwWidth
    w[5] = (uint32_t)(((size_t)size * 16 + 7) / 8);      // dwPitchOrLinearSize
    w[6] = (uint32_t)size;                               // dwDepth
    w[7] = 0u;                                           // dwMipMapCount; w[8..18] dwReserved1[11]
    w[19] = 32u; w[20] = 0x00000004u; w[21] = 0x30315844u;   // ddspf = DDSPF_DX10: size, DDS_FOURCC, 'DX10', zeros
    w[27] = 0x00001000u | 0x00000008u;                   // dwCaps = DDS_SURFACE_FLAGS_TEXTURE | DDS_SURFACE_FLAGS_CUBEMAP
    w[28] = 0x00200000u;                                 // dwCaps2 = DDS_FLAGS_VOLUME; w[29..31] dwCaps3, dwCaps4, dwReserved2
    w[32] = 2u;                                          // header10.dxgiFormat = DXGI_FORMAT_R32G32B32A32_FLOAT
    w[33] = 4u;                                          // resourceDimension = DDS_DIMENSION_TEXTURE3D
    w[34] = 0u;                                          // miscFlag
    w[35] = 1u;                                          // arraySize
    w[36] = 0u;                                          // miscFlags2
    std::memcpy(out, w, 148);
    return 148;
}

int sbx_set_trace_buffer(sbx_ctx* ctx, unsigned long long* dev_records) {
    if (!ctx) return SBX_ERR_INVALID;
    ctx->trace = (CUdeviceptr)(uintptr_t)dev_records;
    return SBX_OK;
}

int sbx_last_timing(sbx_ctx* ctx, sbx_timing* out) {
    if (!ctx || !out) return SBX_ERR_INVALID;
    if (ctx->timing.kernel_ms < 0.0f && ctx->ev1) {
        ctx_scope scope(ctx);
        float ms = 0.0f;
        if (ctx->cu->EventSynchronize(ctx->ev1) == CUDA_SUCCESS &&
            ctx->cu->EventElapsedTime(&ms, ctx->ev0, ctx->ev1) == CUDA_SUCCESS)
            ctx->timing.kernel_ms = ms;
    }
    *out = ctx->timing;
    return SBX_OK;
}

static int op_code(const char* op) {
    static const struct { const char* name; int code; } table[] = {
        {"sinf", 0}, {"cosf", 1}, {"tanf", 2}, {"expf", 3}, {"powf", 4}, {"acosf", 5}, {"atan2f", 6}, {"sqrtf", 7}, {"divf", 8},
        {"hash", 16}, {"hash_arith", 17}, {"noise_iq", 18}, {"noise_w", 19}, {"fbm4", 20}, {"fbm_w3", 21},
        {"sd_sphere", 32}, {"sd_box", 33}, {"sd_torus", 34}, {"sd_y_cylinder", 35}, {"sd_cylinder", 36}, {"sd_bezier", 37},
        {"sd_capsule", 38}, {"sd_plane", 39}, {"op_blend", 40}, {"ik_solver", 41},
        {"henyey_greenstein_phase_func", 48}, {"rayleigh_phase_func", 49}, {"schlick_phase_func", 50},
        {"isotropic_phase_func", 51}, {"fresnel_factor", 52}, {"reflect", 53}, {"refract", 54},
        {"illum_cook_torrance", 55}, {"illum_blinn_phong", 56}, {"intersect_sphere", 57}, {"intersect_plane", 58},
        {"rotate_around_x", 64}, {"rotate_around_y", 65}, {"rotate_around_z", 66}, {"linear_to_srgb", 67}, {"band", 68},
        {"checkboard_pattern", 69}, {"remap", 70}, {"get_primary_ray", 71}, {"smoothstep", 72}, {"mod", 73},
        {"fast_orthonormal_basis", 74}, {"unorm8", 75}};
    for (const auto& e : table)
        if (!std::strcmp(e.name, op)) return e.code;
    return -1;
}

int sbx_eval_op(sbx_ctx* ctx, const char* op, const float* in, int in_stride, float* out, int out_stride, int n) {
    if (!ctx || !op || !in || !out || in_stride <= 0 || out_stride <= 0 || n < 0) return SBX_ERR_INVALID;
    int code = op_code(op);
    if (code < 0) return ctx->fail(SBX_ERR_UNSUPPORTED, "unknown operator %s", op);
    if (n == 0) return SBX_OK;
    ctx_scope scope(ctx);
    int st = ensure_tables(ctx, nullptr);
    if (st != SBX_OK) return st;
    CUdeviceptr din = 0, dout = 0;
    const size_t in_bytes = (size_t)n * in_stride * sizeof(float), out_bytes = (size_t)n * out_stride * sizeof(float);
    SBX_TRY(ctx->cu->MemAlloc(&din, in_bytes), "cuMemAlloc");
    if (ctx->cu->MemAlloc(&dout, out_bytes) != CUDA_SUCCESS) { ctx->cu->MemFree(din); return ctx->fail(SBX_ERR_NOMEM, "cuMemAlloc"); }
    sbx_launch L;
    std::memset(&L, 0, sizeof L);
    sbx_default_params(&L.p, 1, 1);
    L.hash_tab = (const float4*)ctx->hash_tab;
    L.hash_bias = SBX_HASH_MAGIC_BITS + ctx->hash_lo;
    L.hash_len = ctx->hash_len;
    L.hash_span = ctx->hash_len;
    L.lut = (const void*)ctx->lut;
    int rc = SBX_OK;
    do {
        if ((rc = ctx->check(ctx->cu->MemcpyHtoD(din, in, in_bytes), "cuMemcpyHtoD")) != SBX_OK) break;
        if ((rc = ctx->check(ctx->cu->MemcpyHtoD(dout, out, out_bytes), "cuMemcpyHtoD")) != SBX_OK) break;
        void* args[] = {&L, &code, &din, &in_stride, &dout, &out_stride, &n};
        if ((rc = ctx->check(ctx->cu->LaunchKernel(ctx->k_eval, (unsigned)((n + 127) / 128), 1, 1, 128, 1, 1,
                                                   SBX_LUT_MATH_BYTES, nullptr, args, nullptr),
                             "launch sbx_eval_op_kernel")) != SBX_OK) break;
        if ((rc = ctx->check(ctx->cu->MemcpyDtoH(out, dout, out_bytes), "cuMemcpyDtoH")) != SBX_OK) break;
    } while (0);
    ctx->cu->MemFree(din);
    ctx->cu->MemFree(dout);
    return rc;
}

}  // extern "C"
