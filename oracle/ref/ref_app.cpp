// ref_app.cpp -- TEST INFRASTRUCTURE (oracle/_ref).  Never linked into the product library.
//
// Builds the reference's OWN arithmetic: the verbatim headers under /root/reference/src are
// #included where they lie (never copied into this repo) on top of oracle/ref/glsl_shim.h.
// One translation unit per app, selected like the reference selects it (src/Makefile:9):
//     g++ ... -DAPP_CLOUDS -DSBX_REF_HEADER='"app_clouds.h"' -I/root/reference/src
//
// Per-pixel semantics: GLSL / shadertoy per-invocation state.  def.h:7-8 makes file-scope state
// `thread_local`; here the whole app header is wrapped in a struct and a fresh instance is made
// for every pixel, so _mutable state (app_egg.h:188 `depth`, app_atmosphere.h:40 `sun_dir`,
// light.h:14, material.h:17, cornell_box.h:9-12) starts from its initialiser at every pixel --
// the behaviour of every GPU host of the reference (SURVEY.md §8 "Hazards").
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/sbx.h"

#define SBX_CAT2(a, b) a##b
#define SBX_CAT(a, b) SBX_CAT2(a, b)
// one namespace per app: the six translation units end up in one .so and must not share
// (weak, inline) symbols such as app_t::mainImage
#define ref SBX_CAT(ref_, SBX_REF_NAME)

// transcendental call counters (per thread), fed by the GLSL_COUNT hook of the shim
namespace ref {
struct counts_t { unsigned long long sin_, cos_, exp_, pow_, sqrt_, other_; };
static thread_local counts_t tl_counts;
}
#ifdef SBX_REF_NO_COUNT   /* timing build: no counter traffic in the inner loops */
#define GLSL_COUNT(what) ((void)0)
#else
#define GLSL_COUNT(what) (++::ref::tl_counts.what)
#endif
#include "glsl_shim.h"

namespace ref {
using namespace glsl;

// accessors defined BEFORE the reference headers turn u_time / u_mouse into macros
// (uniform_buffer.h:33-35)
static inline float param_time(const sbx_params& p) { return p.u_time; }
static inline vec4 param_mouse(const sbx_params& p) {
    return vec4(p.u_mouse[0], p.u_mouse[1], p.u_mouse[2], p.u_mouse[3]);
}

#define thread_local /* def.h:7-8 -> plain per-instance members */
#define mainImage(a, b) mainImage(vec4& fragColor, const vec2& fragCoord) /* main.h:6-9 `out`/`in` */

struct app_t {
    // what the absent host provides (uniform_buffer.h:32-36)
    vec2 iResolution;
    float iGlobalTime;
    vec4 iMouse;

#include SBX_REF_HEADER

    explicit app_t(const sbx_params& p)
        : iResolution(float(p.width), float(p.height)),
          iGlobalTime(param_time(p)),
          iMouse(param_mouse(p))
#if defined(APP_CLOUDS)
          , wind_dir(p.wind_dir[0], p.wind_dir[1], p.wind_dir[2])
          , sun_dir(p.sun_dir[0], p.sun_dir[1], p.sun_dir[2])
          , sun_color(p.sun_color[0], p.sun_color[1], p.sun_color[2])
          , sun_power(p.sun_power)
          , cld_march_steps(p.cld_march_steps)
          , illum_march_steps(p.illum_march_steps)
          , sigma_scattering(p.sigma_scattering)
          , cld_coverage(p.cld_coverage)
          , cld_thick(p.cld_thick)
          , atm_radius(p.atm_radius)
          , atm_ground_y(p.atm_ground_y)
#elif defined(APP_SDF_AO)
          , fog_density(p.fog_density)
          , fog_falloff(p.fog_falloff)
#endif
    {
    }
};
#undef mainImage
#undef thread_local

static void render_rows(const sbx_params& p, const std::vector<int>& rows, size_t lo, size_t hi,
                        float* out, counts_t* counts) {
    tl_counts = counts_t{};
    for (size_t k = lo; k < hi; ++k) {
        const int y = rows[k];
        float* dst = out + k * size_t(p.width) * 4;
        for (int x = 0; x < p.width; ++x) {
            app_t app(p);  // fresh per-pixel state
            vec4 c;
            app.mainImage(c, vec2(float(x) + 0.5f, float(y) + 0.5f));
            dst[4 * x + 0] = c.x; dst[4 * x + 1] = c.y; dst[4 * x + 2] = c.z; dst[4 * x + 3] = c.w;
        }
    }
    *counts = tl_counts;
}

}  // namespace ref

extern "C" {

// Render the rows of `shard` (NULL = whole frame) compacted into out (rows*width*4 floats) on
// nthreads host threads (rows dealt round-robin in blocks so the load balances).
// counts_out (may be NULL): 7 x uint64 = sin, cos, exp, pow, sqrt, tan/acos/atan, 0 calls over those rows.
int SBX_CAT(sbxref_render_, SBX_REF_NAME)(const sbx_params* p, const sbx_shard* shard, float* out,
                                          int nthreads, unsigned long long* counts_out) {
    if (!p || !out || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    sbx_shard s = {1, 1, 0};
    if (shard && shard->n_parts > 0) s = *shard;
    if (s.stripe_rows <= 0) s.stripe_rows = 1;
    if (s.part < 0 || s.part >= s.n_parts) return SBX_ERR_INVALID;
    std::vector<int> rows;
    for (int y = 0; y < p->height; ++y)
        if ((y / s.stripe_rows) % s.n_parts == s.part) rows.push_back(y);
    if (nthreads <= 0) nthreads = int(std::thread::hardware_concurrency());
    if (nthreads <= 0) nthreads = 1;
    std::atomic<size_t> next(0);
    std::vector<ref::counts_t> counts(size_t(nthreads), ref::counts_t{});
    auto worker = [&](int tid) {
        ref::counts_t total{};
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= rows.size()) break;
            ref::counts_t c{};
            ref::render_rows(*p, rows, k, k + 1, out, &c);
            total.sin_ += c.sin_; total.cos_ += c.cos_; total.exp_ += c.exp_;
            total.pow_ += c.pow_; total.sqrt_ += c.sqrt_; total.other_ += c.other_;
        }
        counts[size_t(tid)] = total;
    };
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    if (counts_out) {
        std::memset(counts_out, 0, 7 * sizeof(unsigned long long));
        for (auto& c : counts) {
            counts_out[0] += c.sin_; counts_out[1] += c.cos_; counts_out[2] += c.exp_;
            counts_out[3] += c.pow_; counts_out[4] += c.sqrt_; counts_out[5] += c.other_;
        }
    }
    return SBX_OK;
}

int SBX_CAT(sbxref_state_bytes_, SBX_REF_NAME)(void) { return int(sizeof(ref::app_t)); }

}  // extern "C"
