// app_clouds_native.h -- hand-written sm_100a version of APP_CLOUDS (src/app_clouds.h), the app the
// headline metric is quoted on.  Same plugin contract as a shaderbox app header (setup_camera /
// setup_scene / render / FOV, then main.h), written directly against the device operator library;
// built by nvcc into images/APP_CLOUDS.native.cubin.  The frame is BIT-IDENTICAL to the unchanged
// reference header compiled as a plugin (tests/test_gpu_parity.py compares both with the oracle):
// every value that reaches the pixel is produced by the reference's operations in the reference's
// order.  What is hand-tuned is which of those operations are executed at all (ncu, profiles/r01b:
// the kernel is FP32-issue bound, 97 % of issued instructions are the fbm octaves):
//
//  * lazy octaves in the view march (density_func :62-86).  density = shape * smoothstep(cov,
//    cov + .0135, shape) is exactly +0 whenever shape <= cov, and the fbm octaves still to come can
//    add at most their gains (noise_iq <= 1): after octave i, t_i + sum(remaining H) <= cov (minus a
//    1e-5 guard, orders of magnitude above the few ulps the fp32 sums can gain) proves the result
//    is 0 without evaluating them.  Half of the sky is empty (cld_coverage .535).
//  * z-slice reuse in the light march (illuminate_volume :107-113).  noise_iq interpolates x, then
//    y, then z (src/noise_iq.h:19-23): the two bilinear z-slice values y0, y1 of an octave depend
//    only on (p.x, p.y, floor(p.z)).  The light march steps by L*dt; when that leaves x and y
//    bit-unchanged (the default sun_dir (0,0,-1) of src/uniform_buffer.h:42) and z stays in the
//    same lattice cell, an octave is one floor, one smoothstep weight and one mix.  Equality is
//    checked on the bits, per sample and per octave, so any sun direction stays exact; it just
//    falls back to the full octave.
//  * exp(-0 * sigma * dt) == 1 exactly: empty light-march samples skip the exponential (:110-113).
//  * the Henyey-Greenstein factor of illuminate_volume (:121) depends only on the ray, not on the
//    sample: one pow per ray instead of one per in-cloud step.
//  * table misses are deferred: lattice indices are clamped into the memo table and the largest
//    one is tracked; only a pixel that ever went out of range is recomputed, on the generic path
//    (no calls and no miss branches inside the march loops).
#include "def.h"
#include "util.h"
#include "intersect.h"

#define hg_g (.2f)
#include "volumetric.h"
#include "noise_iq.h"
#include "fbm.h"

#define cld_noise_factor .001f
DECL_FBM_FUNC(fbm, 4, noise_iq(p))   // :59, the generic form: used when a lattice index leaves the memo table

float sbx_phase;       // henyey_greenstein_phase_func(clamp(dot(L, V), 0, 1)) of :121, per ray
float sbx_cov;         // 1 - cld_coverage (:83)
// z-slice memo of the last fully evaluated sample: position x/y it was taken at and, per octave,
// the lattice z and the two bilinear slice values
float sbx_mx, sbx_my;
float sbx_mz0, sbx_mz1, sbx_mz2, sbx_mz3;
float sbx_a0, sbx_a1, sbx_a2, sbx_a3;   // y0 of octave 0..3
float sbx_b0, sbx_b1, sbx_b2, sbx_b3;   // y1 of octave 0..3
unsigned sbx_kmax;     // largest table index this pixel asked for (see render)

SBX_FN void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) {   // :23-30
    eye = vec3(0.0f, -.5f, 0.0f);
    const float angle = u_mouse.x * .5f;
    look_at = mul(rotate_around_y(angle), vec3(0.0f, 0.0f, -1.0f));
}

SBX_FN void setup_scene() {}

SBX_FN vec3 render_sky_color(_in(vec3) eye_dir) {   // :36-46
    const float sun_amount = max(dot(eye_dir, sun_dir), 0.0f);
    vec3 sky = mix(vec3(.0f, .1f, .4f), vec3(.3f, .6f, .8f), 1.0f - eye_dir.y);
    sky += sun_color * min(pow(sun_amount, 1500.0f) * 5.0f, 1.0f);
    sky += sun_color * min(pow(sun_amount, 10.0f) * .6f, 1.0f);
    return abs(sky);
}

SBX_FN float sbx_weight(float f) { return f * f * __fmaf_rn(f, -2.0f, 3.0f); }   // f*f*(3 - 2f), noise_iq.h:16

// one octave of noise_iq (src/noise_iq.h:11-23) evaluated in full; returns its two z-slices too.
// The lattice index is clamped into the memo table and the largest one remembered: a pixel that
// ever asked for an entry outside the table is recomputed on the generic path (render()).
SBX_FN float octave_full(_in(vec3) x, float& mz, float& y0, float& y1) {
    const vec3 p = floor(x);
    const float wx = sbx_weight(x.x - p.x), wy = sbx_weight(x.y - p.y), wz = sbx_weight(x.z - p.z);
    const float n = p.x + p.y * 157.0f + 113.0f * p.z;
    const unsigned k = (unsigned)(__float_as_int(n + 12582912.0f) - sbx_L->hash_bias);
    sbx_kmax = ::max(sbx_kmax, k);
    const float4* __restrict__ e = sbx_L->hash_tab + ::min(k, (unsigned)sbx_L->hash_span - 1u);
    const float4 z0 = __ldg(e), z1 = __ldg(e + 113);
    y0 = mix(mix(z0.x, z0.y, wx), mix(z0.z, z0.w, wx), wy);
    y1 = mix(mix(z1.x, z1.y, wx), mix(z1.z, z1.w, wx), wy);
    mz = p.z;
    return mix(y0, y1, wz);
}

SBX_FN float sbx_density_of(float shape) { return shape * smoothstep(sbx_cov, sbx_cov + .0135f, shape); }   // :83-84

// density_func (:62-86) for a view-march sample: fbm(pos * 2.03, 2.64, .5, .5) of src/fbm.h:6
// unrolled (H = .5 .25 .125 .0625) with the lazy-octave exits
SBX_FN float density_view(_in(vec3) pos_in) {
    const vec3 pos = pos_in * cld_noise_factor;
    vec3 p = pos * 2.03f;
    const float guard = 1e-5f;
    sbx_mx = pos_in.x; sbx_my = pos_in.y;
    float t = octave_full(p, sbx_mz0, sbx_a0, sbx_b0) * .5f;      // 0 + n*.5 == n*.5 (n >= +0)
    if (t <= sbx_cov - .4375f - guard) return 0.0f;
    p *= 2.64f;
    t += octave_full(p, sbx_mz1, sbx_a1, sbx_b1) * .25f;
    if (t <= sbx_cov - .1875f - guard) return 0.0f;
    p *= 2.64f;
    t += octave_full(p, sbx_mz2, sbx_a2, sbx_b2) * .125f;
    if (t <= sbx_cov - .0625f - guard) return 0.0f;
    p *= 2.64f;
    t += octave_full(p, sbx_mz3, sbx_a3, sbx_b3) * .0625f;
    return sbx_density_of(t);
}

// density_func for a light-march sample.  If x and y are bit-identical to the memoised sample and
// every octave stays in its lattice cell along z, each octave is one weight and one mix.
SBX_FN float density_light(_in(vec3) pos_in) {
    const float z0 = pos_in.z * cld_noise_factor * 2.03f;         // the z chain of p = pos*.001*2.03, p *= 2.64
    const float z1 = z0 * 2.64f, z2 = z1 * 2.64f, z3 = z2 * 2.64f;
    const float c0 = floor(z0), c1 = floor(z1), c2 = floor(z2), c3 = floor(z3);
    float t;
    if (pos_in.x == sbx_mx && pos_in.y == sbx_my && c0 == sbx_mz0 && c1 == sbx_mz1 && c2 == sbx_mz2 && c3 == sbx_mz3) {
        t = mix(sbx_a0, sbx_b0, sbx_weight(z0 - c0)) * .5f;
        t += mix(sbx_a1, sbx_b1, sbx_weight(z1 - c1)) * .25f;
        t += mix(sbx_a2, sbx_b2, sbx_weight(z2 - c2)) * .125f;
        t += mix(sbx_a3, sbx_b3, sbx_weight(z3 - c3)) * .0625f;
    } else {
        const vec3 pos = pos_in * cld_noise_factor;
        vec3 p = pos * 2.03f;
        sbx_mx = pos_in.x; sbx_my = pos_in.y;
        t = octave_full(p, sbx_mz0, sbx_a0, sbx_b0) * .5f;
        p *= 2.64f;
        t += octave_full(p, sbx_mz1, sbx_a1, sbx_b1) * .25f;
        p *= 2.64f;
        t += octave_full(p, sbx_mz2, sbx_a2, sbx_b2) * .125f;
        p *= 2.64f;
        t += octave_full(p, sbx_mz3, sbx_a3, sbx_b3) * .0625f;
    }
    return sbx_density_of(t);
}

SBX_FN float illuminate_volume(_in(vec3) origin, _in(vec3) L) {   // :91-123
    const float dt = cld_thick / float(cld_march_steps);
    vec3 pos = origin;
    float transmittance = 1.0f;
    pos += L * dt;   // don't sample just where the main raymarcher is
    for (int i = 0; i < illum_march_steps; i++) {
        const float density = density_light(pos);
        if (density != 0.0f) transmittance *= exp(-density * sigma_scattering * dt);   // exp(-0) == 1
        pos += L * dt;
    }
    return transmittance * sun_power * sbx_phase;
}

SBX_FN vec4 render_clouds(_in(ray_t) eye) {   // :153-202
    const vec3 projection = eye.direction / eye.direction.y;
    vec3 origin = eye.origin + projection * 150.0f;
    origin += wind_dir * u_time * (1.0f / cld_noise_factor);

    sbx_cov = 1.0f - cld_coverage;
    sbx_phase = henyey_greenstein_phase_func(clamp(dot(sun_dir, eye.direction), 0.0f, 1.0f));

    volume_sampler_t cloud = construct_volume(origin);
    float t = 0.0f;
    const float dt = cld_thick / float(cld_march_steps);
    for (int i = 0; i < cld_march_steps; i++) {
        cloud.pos = cloud.origin + t * projection;
        t += dt;
        const float density = density_view(cloud.pos);
        if (!(density < .005f)) {                                          // integrate_volume, :125-148
            const float T_i = exp(-density * sigma_scattering * dt);       // Beer-Lambert
            cloud.transmittance *= T_i;
            cloud.radiance += (density * sigma_scattering) * illuminate_volume(cloud.pos, sun_dir) *
                              cloud.transmittance * dt;
            cloud.alpha += (1.0f - T_i) * (1.0f - cloud.alpha);
        }
        if (cloud.alpha > .999f) break;
    }
    const float cutoff = dot(eye.direction, vec3(0.0f, 1.0f, 0.0f));
    return vec4(cloud.radiance, cloud.alpha * smoothstep(.0f, .2f, cutoff));
}

// ---- the generic path: the app as written (:62-202) on the library's noise_iq / fbm, which fall
// back to the arithmetic hash outside the memo table.  Cold code: runs only for a pixel whose fast
// path met a lattice index outside the table (huge u_time * wind_dir, or the table switched off).
SBX_FN float generic_density(_in(vec3) pos_in) {
    const vec3 pos = pos_in * cld_noise_factor;
    return sbx_density_of(fbm(pos * 2.03f, 2.64f, .5f, .5f));
}
SBX_FN vec4 generic_clouds(_in(ray_t) eye) {
    const vec3 projection = eye.direction / eye.direction.y;
    vec3 origin = eye.origin + projection * 150.0f;
    origin += wind_dir * u_time * (1.0f / cld_noise_factor);
    volume_sampler_t cloud = construct_volume(origin);
    float t = 0.0f;
    const float dt = cld_thick / float(cld_march_steps);
    for (int i = 0; i < cld_march_steps; i++) {
        cloud.pos = cloud.origin + t * projection;
        t += dt;
        const float density = generic_density(cloud.pos);
        if (!(density < .005f)) {
            const float T_i = exp(-density * sigma_scattering * dt);
            cloud.transmittance *= T_i;
            vec3 lp = cloud.pos;
            float tr = 1.0f;
            lp += sun_dir * dt;
            for (int j = 0; j < illum_march_steps; j++) {
                tr *= exp(-generic_density(lp) * sigma_scattering * dt);
                lp += sun_dir * dt;
            }
            cloud.radiance += (density * sigma_scattering) * (tr * sun_power * sbx_phase) * cloud.transmittance * dt;
            cloud.alpha += (1.0f - T_i) * (1.0f - cloud.alpha);
        }
        if (cloud.alpha > .999f) break;
    }
    const float cutoff = dot(eye.direction, vec3(0.0f, 1.0f, 0.0f));
    return vec4(cloud.radiance, cloud.alpha * smoothstep(.0f, .2f, cutoff));
}
// out of line, on a FRESH app object, so that the hot path's state never has its address taken
static __device__ __noinline__ float4 sbx_generic_pixel(const sbx_launch* L, float ox, float oy, float oz, float dx,
                                                        float dy, float dz) {
    sbx_app a(L);
    a.sbx_cov = 1.0f - a.cld_coverage;
    ray_t eye;
    eye.origin = vec3(ox, oy, oz);
    eye.direction = vec3(dx, dy, dz);
    a.sbx_phase = a.henyey_greenstein_phase_func(clamp(dot(a.sun_dir, eye.direction), 0.0f, 1.0f));
    const vec4 c = a.generic_clouds(eye);
    return make_float4(c.x, c.y, c.z, c.w);
}

SBX_FN vec3 render(_in(ray_t) eye_ray, _in(vec3) point_cam) {   // :204-218
    const vec3 sky = render_sky_color(eye_ray.direction);
    if (dot(eye_ray.direction, vec3(0.0f, 1.0f, 0.0f)) < 0.05f) return sky;
    sbx_kmax = 0u;
    vec4 cld = render_clouds(eye_ray);
    if (sbx_kmax >= (unsigned)sbx_L->hash_span) {                       // table miss: redo the pixel
        const float4 g = sbx_generic_pixel(sbx_L, eye_ray.origin.x, eye_ray.origin.y, eye_ray.origin.z,
                                           eye_ray.direction.x, eye_ray.direction.y, eye_ray.direction.z);
        cld = vec4(g.x, g.y, g.z, g.w);
    }
    const vec3 col = mix(sky, cld.rgb, cld.a);
    return abs(col);
}

#define FOV 1.0f   // :220
#include "main.h"
