// sbx_kernel.cuh -- the pixel loop as a CUDA grid (replaces the host loops of the reference:
// VML's SDL_app.cpp, src/Makefile:21, and hlsltoy's fullscreen triangle,
// util/hlsltoy/src/hlsltoy.cpp:494-495).
//
// Translation unit layout (assembled by sbx_compile_app with NVRTC, -default-device):
//     -DAPP_xxx -DSBX_APP_HEADER="<the app header, float literals suffixed>"
//     #include "sbx/sbx_kernel.cuh"
// The UNCHANGED app header is included inside `struct sbx_app`, so its file-scope state becomes
// per-pixel members and its free functions become __device__ members; the library headers it
// pulls in ("def.h", "sdf.h", "noise_iq.h", ...) resolve to this directory.
//
// Mapping: one thread per pixel; a warp owns an 8x4 pixel tile (each tile row is one 128-byte
// line of the RGBA32F frame, so a warp's store is four full lines); a CTA is SBX_WARPS_PER_CTA
// warps on consecutive tiles.  A scene kernel may instead put SBX_LANES_PER_PIXEL = P > 1 lanes on
// every pixel (native/app_clouds_native.h splits a ray's march steps over them): the warp tile is
// then (32/P) x 1 pixels, lane = pixel * P + phase, and only phase 0 stores.  Tiles are issued
// bottom row first (measured: issuing the top rows first costs APP_CLOUDS 6 % -- its longest rays
// are the ones just above the horizon, and they should not start last).  The prologue stages the
// math LUT block from HBM to shared memory with one TMA bulk copy (cp.async.bulk) completing on
// an mbarrier.
#ifndef SBX_KERNEL_CUH_
#define SBX_KERNEL_CUH_

#include "sbx_launch.h"
#include "sbx_vec.cuh"

#ifndef SBX_WARPS_PER_CTA
#define SBX_WARPS_PER_CTA 4
#endif
#ifndef SBX_MIN_CTAS_PER_SM
#define SBX_MIN_CTAS_PER_SM 1
#endif
#ifndef SBX_LANES_PER_PIXEL
#define SBX_LANES_PER_PIXEL 1
#endif
#ifndef SBX_HYBRID_LANES
#define SBX_HYBRID_LANES 1
#endif
#if SBX_HYBRID_LANES > 1 && SBX_LANES_PER_PIXEL > 1
#error "a hybrid image marches its first region with one lane per pixel"
#endif
#if SBX_LANES_PER_PIXEL == 1
#define SBX_IMG_TILE_W SBX_TILE_W
#define SBX_IMG_TILE_H SBX_TILE_H
#else
#define SBX_IMG_TILE_W (32 / SBX_LANES_PER_PIXEL)
#define SBX_IMG_TILE_H 1
#endif
// read by the host at load time (cuModuleGetGlobal): { tile width, tile height, lanes per pixel, lanes per pixel of the
// second region of a hybrid image (0 = not hybrid) }
extern "C" __device__ const int sbx_image_info[4] = {SBX_IMG_TILE_W, SBX_IMG_TILE_H, SBX_LANES_PER_PIXEL,
                                                     SBX_HYBRID_LANES > 1 ? SBX_HYBRID_LANES : 0};

// read by the host at load time: { per-mille of the frame height, from the bottom, whose pixels are known to be trivial
// (they return before any march) -- the launch issues those rows LAST; 1 if the image needs the noise textures; 0, 0 }.  A scene kernel sets SBX_HINT_TRIVIAL_ROWS.
#ifndef SBX_HINT_TRIVIAL_ROWS
#define SBX_HINT_TRIVIAL_ROWS 0
#endif
extern "C" __device__ const int sbx_image_hints[4] = {SBX_HINT_TRIVIAL_ROWS,
#ifdef SBX_USES_NOISE_TEX
                                                      1,   // [1]: the kernel takes a second parameter, sbx_tex_params
#else
                                                      0,
#endif
                                                      0, 0};

// ---- TMA bulk copy of the LUT block: global -> shared, completion on an mbarrier --------------
__device__ __forceinline__ unsigned sbx_smem_addr(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sbx_stage_lut(const void* lut_global) {
    __shared__ __align__(8) unsigned long long sbx_lut_bar;
    if (threadIdx.x == 0) {
        // one thread arms the barrier, issues the bulk copy and waits for its bytes to land; the others
        // just park on the CTA barrier below (a spinning try_wait in every thread cost cheap apps 4 % of
        // their issue slots -- profiles/r01g)
        const unsigned bar = sbx_smem_addr(&sbx_lut_bar);
        const unsigned dst = sbx_smem_addr(sbx_smem);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(SBX_LUT_MATH_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(dst), "l"(lut_global), "n"(SBX_LUT_MATH_BYTES), "r"(bar) : "memory");
        unsigned done = 0;
        while (!done) {   // phase 0 completes when the 512 bytes have been written (acquire)
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n" : "=r"(done) : "r"(bar) : "memory");
        }
    }
    __syncthreads();   // orders the table (observed by thread 0) before every other thread's reads
}

namespace sbx_glsl {

// uniform loaders used by uniform_buffer.h's SBX_UNIFORM()
SBX_FN float sbx_uniform(const float& v) { return v; }
SBX_FN int sbx_uniform(const int& v) { return v; }
SBX_FN vec3 sbx_uniform(const float (&v)[3]) { return vec3(v[0], v[1], v[2]); }
// read before uniform_buffer.h turns u_time / u_mouse into macros
SBX_FN float sbx_param_time(const sbx_launch* L) { return L->times ? __ldg(L->times + blockIdx.y) : L->p.u_time; }
SBX_FN vec4 sbx_param_mouse(const sbx_launch* L) {
    return vec4(L->p.u_mouse[0], L->p.u_mouse[1], L->p.u_mouse[2], L->p.u_mouse[3]);
}

struct sbx_app {
    const sbx_launch* __restrict__ sbx_L;   // must stay the first member (see uniform_buffer.h)
    // what the host provides to the C++ build of the reference (src/uniform_buffer.h:32-36)
    vec2 iResolution;
    float iGlobalTime;
    vec4 iMouse;
    bool sbx_coop;                          // this warp marches with several lanes per pixel (warp-uniform)
    const sbx_tex_params* __restrict__ sbx_T;   // the 3-D noise textures (images built with -DSBX_USES_NOISE_TEX), else NULL
    unsigned sbx_lanes;                     // the lanes of this warp that render a pixel (ballot taken with all 32 present)

#include SBX_APP_HEADER

    __device__ __forceinline__ explicit sbx_app(const sbx_launch* L, bool coop = false, const sbx_tex_params* T = nullptr,
                                                unsigned lanes = 0xffffffffu)
        : sbx_L(L),
          iResolution(float(L->p.width), float(L->p.height)),
          iGlobalTime(sbx_param_time(L)),
          iMouse(sbx_param_mouse(L)),
          sbx_coop(coop),
          sbx_T(T),
          sbx_lanes(lanes) {}
};

}  // namespace sbx_glsl

// The last CTA to finish publishes done_value at done_flag (own HBM, a peer GPU's, or mapped host memory).  Every
// thread's pixel stores are ordered before the CTA barrier; thread 0's system-scope fence is cumulative over them;
// the counter hands "all CTAs have fenced" to the last one, which fences again before the flag store, so whoever
// acquires the flag (cuStreamWaitValue32, a polling kernel, a host read) sees every pixel of this launch.
__device__ __forceinline__ void sbx_signal_done(const sbx_launch& L) {
    if (L.done_flag == nullptr) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned ctas = gridDim.x * gridDim.y;
        if (atomicAdd(L.done_counter, 1u) + 1u == ctas) {
            *reinterpret_cast<volatile unsigned*>(L.done_counter) = 0u;   // ready for the next launch on this stream
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(L.done_flag), "r"(L.done_value) : "memory");
        }
    }
}

// The first SBX_PREFETCH_CTAS thread blocks of a launch each ask the L2 for one slice of the lattice-hash memo table
// (cp.async.bulk.prefetch.L2: one instruction, no completion to wait for).  After an L2 flush -- or on the first frame --
// the table's lines would otherwise arrive one dependent miss at a time (octave after octave, ~0.7 us each from HBM);
// the whole 8 MB is 1.3 us of HBM bandwidth.
#ifndef SBX_PREFETCH_CTAS
#define SBX_PREFETCH_CTAS 128
#endif
__device__ __forceinline__ void sbx_prefetch_tables(const sbx_launch& L) {
#if SBX_PREFETCH_CTAS > 0
    if (threadIdx.x == 32 && blockIdx.x < SBX_PREFETCH_CTAS && blockIdx.y == 0 && L.hash_len > 0) {
        const unsigned long long total = (unsigned long long)L.hash_len * 32ull;               // bytes (a multiple of 16 KB)
        const unsigned long long slice = (total / SBX_PREFETCH_CTAS + 15ull) & ~15ull;
        const unsigned long long at = slice * blockIdx.x;
        if (at < total) {
            const unsigned bytes = (unsigned)(total - at < slice ? total - at : slice);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const char*>(L.hash_tab) + at), "r"(bytes) : "memory");
        }
    }
#endif
}

#ifdef SBX_USES_NOISE_TEX
#define SBX_TEX_PARAM , const __grid_constant__ sbx_tex_params T
#define SBX_TEX_ARG &T
#else
#define SBX_TEX_PARAM
#define SBX_TEX_ARG nullptr
#endif
extern "C" __global__ void __launch_bounds__(SBX_WARPS_PER_CTA * 32, SBX_MIN_CTAS_PER_SM)
sbx_render(const __grid_constant__ sbx_launch L SBX_TEX_PARAM) {
    sbx_prefetch_tables(L);
    sbx_stage_lut(L.lut);

    int warp = blockIdx.x * SBX_WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
#ifdef SBX_TRACE
    const int trace_warp = warp;
    unsigned long long trace_t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(trace_t0));
#endif
#if SBX_HYBRID_LANES > 1
    // hybrid image: the first reg[0].warps warps render 8x4 tiles with one lane per pixel, the warps after them
    // render the remaining (upper) rows with SBX_HYBRID_LANES lanes per pixel -- the tail of the launch is cut
    // into pieces P times shorter, so it drains P times faster, and only the tail pays the cooperative overhead
    const bool coop = warp >= L.reg[0].warps;
    if (coop) warp -= L.reg[0].warps;
    const sbx_region& R = L.reg[coop ? 1 : 0];
    const int P = coop ? SBX_HYBRID_LANES : 1;
#else
    const bool coop = SBX_LANES_PER_PIXEL > 1;
    const sbx_region& R = L.reg[0];
    const int P = SBX_LANES_PER_PIXEL;
#endif
    const int tile_w = P == 1 ? SBX_TILE_W : 32 / P, tile_h = P == 1 ? SBX_TILE_H : 1;
    // warp -> (tile row, column slot): division by the host's magic number (0 = divide)
    int trow = R.magic ? (int)(((unsigned long long)(unsigned)warp * R.magic) >> 40) : warp / R.tiles_per_row;
    int tile_x = warp - trow * R.tiles_per_row;
    // issue order: tile rows from first_tile_row upwards, then the ones below it (an app whose bottom rows are trivial --
    // APP_CLOUDS' sky under the horizon -- says so in sbx_image_hints, and those rows then fill the end of the launch)
    if (warp < R.warps) {
        trow += R.first_tile_row;
        trow -= trow >= R.tile_rows ? R.tile_rows : 0;
    }
    const int lr0 = R.row0 + trow * tile_h;
    if (L.col_parts > 1) {   // this part's columns of the row: (tile_x + lr0 / 4) % col_parts == col_part
        int first = (L.col_part - (lr0 >> 2)) % L.col_parts;
        first += first < 0 ? L.col_parts : 0;
        tile_x = first + tile_x * L.col_parts;
    }
    int x, lr;
    bool valid;
    if (P == 1) {
        x = tile_x * SBX_TILE_W + (lane & (SBX_TILE_W - 1));
        lr = lr0 + (lane / SBX_TILE_W);
        valid = warp < R.warps && x < L.p.width && lr < R.row0 + R.rows;
    } else {
        // P lanes per pixel cooperate through warp shuffles, so every lane of the warp stays: lanes past
        // the right edge (or in a tile past the last one) render the clamped pixel and store nothing
        x = tile_x * tile_w + lane / P;
        lr = lr0;
        valid = warp < R.warps && x < L.p.width && lr < R.row0 + R.rows && (lane % P) == 0;
        x = x < L.p.width ? x : L.p.width - 1;
        lr = lr < L.local_rows ? lr : L.local_rows - 1;
    }

    const unsigned lanes = __ballot_sync(0xffffffffu, valid || P > 1);   // who enters the pixel body (all 32 lanes vote here)
    if (valid || P > 1) {
        // local (compacted) row -> frame row of this shard
        const int y = L.n_parts == 1 ? lr : ((lr / L.stripe_rows) * L.n_parts + L.part) * L.stripe_rows + lr % L.stripe_rows;

        sbx_glsl::sbx_app app(&L, coop, SBX_TEX_ARG, lanes);
        sbx_glsl::vec4 c;
        app.mainImage(c, sbx_glsl::vec2(float(x) + 0.5f, float(y) + 0.5f));

        if (valid) {
            const size_t at = ((size_t)blockIdx.y * (size_t)L.local_rows + (size_t)(L.out_is_frame ? y : lr)) * (size_t)L.p.width + (size_t)x;
            if (L.out_rgba8) __stcs(reinterpret_cast<unsigned*>(L.out) + at, sbx_pack_unorm8(c.x, c.y, c.z, c.w));
            else __stcs(reinterpret_cast<float4*>(L.out) + at, make_float4(c.x, c.y, c.z, c.w));
        }
    }
#ifdef SBX_TRACE
    __syncwarp();
    if (L.trace != nullptr && lane == 0) {
        unsigned long long t1;
        unsigned sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        unsigned long long* rec = L.trace + 4ull * (unsigned long long)trace_warp;
        rec[0] = trace_t0; rec[1] = t1; rec[2] = sm; rec[3] = coop ? 1u : 0u;
    }
#endif
    sbx_signal_done(L);
}

#endif  // SBX_KERNEL_CUH_
