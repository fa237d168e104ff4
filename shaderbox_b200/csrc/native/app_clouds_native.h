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
//    one is tracked; only if it was out of range is the sample recomputed on the arithmetic path.
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
float sbx_mz[4], sbx_y0[4], sbx_y1[4];
unsigned sbx_kmax;     // largest (clamped) table index seen by the current sample

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

// one octave of noise_iq (src/noise_iq.h:11-23) evaluated in full; leaves its z-slices in the memo
SBX_FN float octave_full(_in(vec3) x, const int o) {
    const vec3 p = floor(x);
    const float wx = sbx_weight(x.x - p.x), wy = sbx_weight(x.y - p.y), wz = sbx_weight(x.z - p.z);
    const float n = p.x + p.y * 157.0f + 113.0f * p.z;
    const unsigned k = (unsigned)(__float_as_int(n + 12582912.0f) - sbx_L->hash_bias);
    sbx_kmax = ::max(sbx_kmax, k);
    const float4* __restrict__ e = sbx_L->hash_tab + ::min(k, (unsigned)sbx_L->hash_span - 1u);
    const float4 z0 = __ldg(e), z1 = __ldg(e + 113);
    const float y0 = mix(mix(z0.x, z0.y, wx), mix(z0.z, z0.w, wx), wy);
    const float y1 = mix(mix(z1.x, z1.y, wx), mix(z1.z, z1.w, wx), wy);
    sbx_mz[o] = p.z; sbx_y0[o] = y0; sbx_y1[o] = y1;
    return mix(y0, y1, wz);
}

// the same octave when x.x, x.y are bit-identical to the memoised sample: only z is new
SBX_FN float octave_z(_in(vec3) x, const int o) {
    const float pz = floor(x.z);
    if (pz != sbx_mz[o]) return octave_full(x, o);
    return mix(sbx_y0[o], sbx_y1[o], sbx_weight(x.z - pz));
}

SBX_FN float sbx_density_of(float shape) { return shape * smoothstep(sbx_cov, sbx_cov + .0135f, shape); }   // :83-84
SBX_FN bool sbx_table_missed() { return sbx_kmax >= (unsigned)sbx_L->hash_span || sbx_L->hash_span <= 0; }
static __device__ __noinline__ float sbx_density_generic(sbx_app* self, float x, float y, float z) {
    const vec3 pos = vec3(x, y, z) * cld_noise_factor;
    return self->sbx_density_of(self->fbm(pos * 2.03f, 2.64f, .5f, .5f));
}

// density_func (:62-86) for a view-march sample: fbm(pos * 2.03, 2.64, .5, .5) of src/fbm.h:6
// unrolled (H = .5 .25 .125 .0625) with the lazy-octave exits
SBX_FN float density_view(_in(vec3) pos_in) {
    const vec3 pos = pos_in * cld_noise_factor;
    vec3 p = pos * 2.03f;
    const float guard = 1e-5f;
    sbx_kmax = 0u;
    sbx_mx = pos_in.x; sbx_my = pos_in.y;
    float t = octave_full(p, 0) * .5f;                   // 0 + n*.5 == n*.5 (n >= +0)
    if (t <= sbx_cov - .4375f - guard) { sbx_mx = __int_as_float(0x7fc00000); return sbx_table_missed() ? sbx_density_generic(this, pos_in.x, pos_in.y, pos_in.z) : 0.0f; }
    p *= 2.64f;
    t += octave_full(p, 1) * .25f;
    if (t <= sbx_cov - .1875f - guard) { sbx_mx = __int_as_float(0x7fc00000); return sbx_table_missed() ? sbx_density_generic(this, pos_in.x, pos_in.y, pos_in.z) : 0.0f; }
    p *= 2.64f;
    t += octave_full(p, 2) * .125f;
    if (t <= sbx_cov - .0625f - guard) { sbx_mx = __int_as_float(0x7fc00000); return sbx_table_missed() ? sbx_density_generic(this, pos_in.x, pos_in.y, pos_in.z) : 0.0f; }
    p *= 2.64f;
    t += octave_full(p, 3) * .0625f;
    if (sbx_table_missed()) { sbx_mx = __int_as_float(0x7fc00000); return sbx_density_generic(this, pos_in.x, pos_in.y, pos_in.z); }
    return sbx_density_of(t);
}

// density_func for a light-march sample: z-slice reuse when x and y are those of the memoised sample
SBX_FN float density_light(_in(vec3) pos_in) {
    const vec3 pos = pos_in * cld_noise_factor;
    vec3 p = pos * 2.03f;
    sbx_kmax = 0u;
    float t;
    if (pos_in.x == sbx_mx && pos_in.y == sbx_my) {
        t = octave_z(p, 0) * .5f;
        p *= 2.64f;
        t += octave_z(p, 1) * .25f;
        p *= 2.64f;
        t += octave_z(p, 2) * .125f;
        p *= 2.64f;
        t += octave_z(p, 3) * .0625f;
    } else {
        sbx_mx = pos_in.x; sbx_my = pos_in.y;
        t = octave_full(p, 0) * .5f;
        p *= 2.64f;
        t += octave_full(p, 1) * .25f;
        p *= 2.64f;
        t += octave_full(p, 2) * .125f;
        p *= 2.64f;
        t += octave_full(p, 3) * .0625f;
    }
    if (sbx_table_missed()) { sbx_mx = __int_as_float(0x7fc00000); return sbx_density_generic(this, pos_in.x, pos_in.y, pos_in.z); }
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

SBX_FN vec3 render(_in(ray_t) eye_ray, _in(vec3) point_cam) {   // :204-218
    const vec3 sky = render_sky_color(eye_ray.direction);
    if (dot(eye_ray.direction, vec3(0.0f, 1.0f, 0.0f)) < 0.05f) return sky;
    const vec4 cld = render_clouds(eye_ray);
    const vec3 col = mix(sky, cld.rgb, cld.a);
    return abs(col);
}

#define FOV 1.0f   // :220
#include "main.h"
