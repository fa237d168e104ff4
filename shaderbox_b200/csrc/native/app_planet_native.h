// app_planet_native.h -- hand-written sm_100a version of APP_PLANET (src/app_planet.h), the 3840x2160
// multi-GPU configuration of BASELINE.json.  Same plugin contract as a shaderbox app header
// (setup_camera / setup_scene / render / FOV, then main.h), written against the device operator
// library; built by nvcc into images/APP_PLANET.native.cubin.  The frame is BIT-IDENTICAL to the
// unchanged reference header compiled as a plugin: every value that reaches the pixel is produced
// by the reference's operations in the reference's order.  What is hand-tuned is which noise
// octaves are evaluated at all (the unchanged header needs 255 registers and ~6 noise_iq calls per
// terrain step, 4 per cloud step, 84 per surface normal):
//
//  * lazy fbm octaves.  Every fbm of the app feeds a smoothstep whose lower edge e0 is a literal
//    (terrain: smoothstep(.35, 1, h0), smoothstep(.6, 1, h1), :177-181; clouds: smoothstep(cov,
//    cov + fuzzy, dens), :117).  smoothstep is exactly +0 for x <= e0, and the octaves still to
//    come add at most their gains (the bases noise, |2 noise - 1| and 1 - |2 noise - 1| all lie in
//    [0, 1]): once t_i + sum(remaining H) <= e0 (minus a 1e-5 guard, orders of magnitude above
//    the few ulps the fp32 sums can gain) the result is 0 without evaluating them.  The ridged
//    terrain term (edge .6, first gain .454) needs its 2nd and 3rd octave for a third of the samples.
//  * clouds_map (:109-125) outside the height band.  dens is multiplied by band(.2, .35, .65, h)
//    = smoothstep(.2,.35,h) * (1 - smoothstep(.35,.65,h)), exactly +0 for h <= .2 (first factor 0)
//    and for h >= .65 (second factor 1 - 1: the quotient of the smoothstep is >= 1 because rounding
//    is monotonic).  With dens == +0 integrate_volume (:83-107) changes nothing: T_i = exp(-0) = 1,
//    radiance += 0, alpha += 0 * (1 - alpha).  So such a step skips its 4-octave fbm and both
//    exponentials; so does a step whose dens is proven <= cld_coverage.
//  * per-ray work after the atmosphere test: the three rotation matrices (:310-312) are only
//    built for rays that hit the atmosphere shell.
#include "def.h"
#include "util.h"
#include "intersect.h"

#define hg_g (.76f)
#include "volumetric.h"
#include "noise_iq.h"
#include "fbm.h"

#define max_height .4f
#define max_ray_dist (max_height * 4.0f)
#define planet_radius 1.0f                       // planet = sphere_t{vec3(0,0,0), 1., 0} (:15-17); x - 0 == x, so the
                                                 // `- planet.origin` of :140,:158,:331 is dropped

// ---- lazy fbm (src/fbm.h:6 with an early "cannot exceed the edge" exit) ------------------------
// Returns false when fbm(pos) is proven <= edge (out is not written); otherwise true and out = the
// reference's fbm value.  `_basis` is an expression over p with values in [0, 1].
#define SBX_LAZY_FBM(_name, _octaves, _basis)                                                         \
    SBX_FN bool _name(_in(vec3) pos, float lacunarity, float init_gain, float gain, float edge, float& out) { \
        float Hs[_octaves], rem[_octaves];                                                            \
        Hs[0] = init_gain;                                                                            \
        _Pragma("unroll") for (int i = 1; i < _octaves; i++) Hs[i] = Hs[i - 1] * gain;                \
        rem[_octaves - 1] = 0.0f;                                                                     \
        _Pragma("unroll") for (int i = _octaves - 2; i >= 0; i--) rem[i] = rem[i + 1] + Hs[i + 1];    \
        vec3 p = pos;                                                                                 \
        float t = 0.0f;                                                                               \
        _Pragma("unroll") for (int i = 0; i < _octaves; i++) {                                        \
            t += _basis * Hs[i];                                                                      \
            p *= lacunarity;                                                                          \
            if (i < _octaves - 1 && t + rem[i] <= edge - 1e-5f) return false;                         \
        }                                                                                             \
        out = t;                                                                                      \
        return true;                                                                                  \
    }

// The same with the octave loop kept as a loop: the 7-octave normal fbms are evaluated 12 times per hit
// pixel, once each; unrolled they are most of a 195 KB kernel that no longer fits the instruction
// cache (ncu: 5.7 "no instruction" stall cycles per issued instruction).  The remaining-gain bound is
// kept by subtraction, which can sit a few ulps under the exact suffix sum -- far inside the guard.
#define SBX_LAZY_FBM_LOOP(_name, _octaves, _basis)                                                    \
    SBX_FN bool _name(_in(vec3) pos, float lacunarity, float init_gain, float gain, float edge, float& out) { \
        float rem = 0.0f, Hn = init_gain;                                                             \
        _Pragma("unroll") for (int i = 1; i < _octaves; i++) { Hn *= gain; rem += Hn; }               \
        vec3 p = pos;                                                                                 \
        float H = init_gain, t = 0.0f;                                                                \
        _Pragma("unroll 1") for (int i = 0; i < _octaves; i++) {                                      \
            t += _basis * H;                                                                          \
            p *= lacunarity;                                                                          \
            H *= gain;                                                                                \
            if (t + rem <= edge - 1e-5f) { if (i < _octaves - 1) return false; }                      \
            rem -= H;                                                                                 \
        }                                                                                             \
        out = t;                                                                                      \
        return true;                                                                                  \
    }

#define sbx_anoise (abs(noise_iq(p) * 2.0f - 1.0f))              // :67
#define sbx_rnoise (1.0f - abs(noise_iq(p) * 2.0f - 1.0f))       // :168
SBX_LAZY_FBM(fbm_clouds, 4, sbx_anoise)                          // :68
SBX_LAZY_FBM(fbm_terr, 3, noise_iq(p))                           // :170
SBX_LAZY_FBM(fbm_terr_r, 3, sbx_rnoise)                          // :171
SBX_LAZY_FBM_LOOP(fbm_terr_normals, 7, noise_iq(p))              // :173
SBX_LAZY_FBM_LOOP(fbm_terr_r_normals, 7, sbx_rnoise)             // :174

SBX_FN vec3 background(_in(ray_t) eye) {   // :22-42
    const vec3 sun_color = vec3(1.0f, .9f, .55f);
    const float sun_amount = clamp(dot(eye.direction, vec3(0.0f, 0.0f, 1.0f)), 0.0f, 1.0f);
    vec3 sky = mix(vec3(.0f, .05f, .2f), vec3(.15f, .3f, .4f), 1.0f - eye.direction.y);
    sky += sun_color * clamp(pow(sun_amount, 30.0f) * 5.0f, 0.0f, 1.0f);
    sky += sun_color * clamp(pow(sun_amount, 10.0f) * .6f, 0.0f, 1.0f);
    return abs(sky);
}

SBX_FN void setup_scene() {}

SBX_FN void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) {   // :48-59
    eye = vec3(0.0f, 0.0f, -2.5f);
    look_at = vec3(0.0f, 0.0f, 2.0f);
}

// ---- clouds ------------------------------------------------------------------------------------
#define vol_coeff_absorb 30.034f
#define cld_coverage_k .29475675f   // higher = less clouds (:113)
#define cld_fuzzy .0335f            // (:114)
volume_sampler_t cloud;             // :71

// clouds_map + integrate_volume (:83-125) for the sample at cloud.pos / cloud.height
SBX_FN void clouds_map(float t_step) {
    if (cloud.height <= .2f || cloud.height >= .65f) return;         // band(.2, .35, .65, h) == +0: the step integrates nothing
    float dens;
    if (!fbm_clouds(cloud.pos * 3.2343f + vec3(.35f, 13.35f, 2.67f), 2.0276f, .5f, .5f, cld_coverage_k, dens)) return;
    dens *= smoothstep(cld_coverage_k, cld_coverage_k + cld_fuzzy, dens);
    dens *= band(.2f, .35f, .65f, cloud.height);
    if (dens == 0.0f) return;                                        // (+0: see the header comment)
    // integrate_volume (:83-107); illuminate_volume (:73-81) is exp(height) / .055
    const float T_i = exp(-vol_coeff_absorb * dens * t_step);
    cloud.transmittance *= T_i;
    cloud.radiance += dens * (exp(cloud.height) / .055f) * cloud.transmittance * t_step;
    cloud.alpha += (1.0f - T_i) * (1.0f - cloud.alpha);
}

SBX_FN void clouds_march(_in(ray_t) eye, float max_travel, _in(mat3) rot) {   // :127-147
    const int steps = 75;
    const float t_step = max_ray_dist / float(steps);
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        if (t > max_travel || cloud.alpha >= 1.0f) return;
        const vec3 o = cloud.origin + t * eye.direction;
        cloud.pos = mul(rot, o);
        cloud.height = (length(cloud.pos) - planet_radius) / max_height;
        t += t_step;
        clouds_map(t_step);
    }
}

SBX_FN void clouds_shadow_march(_in(vec3) dir, _in(mat3) rot) {   // :149-166
    const int steps = 5;
    const float t_step = max_height / float(steps);
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        const vec3 o = cloud.origin + t * dir;
        cloud.pos = mul(rot, o);
        cloud.height = (length(cloud.pos) - planet_radius) / max_height;
        t += t_step;
        clouds_map(t_step);
    }
}

// ---- terrain -----------------------------------------------------------------------------------
#define TERR_STEPS 120
#define TERR_EPS .005f

// (distance, height) of :175-187; n0 / n1 are exactly +0 when their fbm stays below the smoothstep edge
SBX_FN vec2 sdf_terrain_map(_in(vec3) pos) {
    float h, n0 = 0.0f, n1 = 0.0f;
    if (fbm_terr(pos * 2.0987f, 2.0244f, .454f, .454f, .35f, h)) n0 = smoothstep(.35f, 1.0f, h);
    if (fbm_terr_r(pos * 1.50987f + vec3(1.9489f, 2.435f, .5483f), 2.0244f, .454f, .454f, .6f, h)) n1 = smoothstep(.6f, 1.0f, h);
    const float n = n0 + n1;
    return vec2(length(pos) - planet_radius - n * max_height, n / max_height);
}

SBX_FN float sdf_terrain_map_detail(_in(vec3) pos) {   // :189-201, .x only (the normal is its only user)
    float h, n0 = 0.0f, n1 = 0.0f;
    if (fbm_terr_normals(pos * 2.0987f, 2.0244f, .454f, .454f, .35f, h)) n0 = smoothstep(.35f, 1.0f, h);
    if (fbm_terr_r_normals(pos * 1.50987f + vec3(1.9489f, 2.435f, .5483f), 2.0244f, .454f, .454f, .6f, h)) n1 = smoothstep(.6f, 1.0f, h);
    const float n = n0 + n1;
    return length(pos) - planet_radius - n * max_height;
}

SBX_FN vec3 sdf_terrain_normal(_in(vec3) p) {   // :203-214: central differences, dt = (0.001, 0, 0) swizzled
    // F(p + dt.xzz) - F(p - dt.xzz) etc.: the offset vector is added / subtracted whole, zeros included
    // (x + 0 turns -0 into +0, like the reference).  One loop body for the three axes (code size, see above).
    float f[3];
#pragma unroll 1
    for (int axis = 0; axis < 3; axis++) {
        const vec3 d = vec3(axis == 0 ? 0.001f : 0.0f, axis == 1 ? 0.001f : 0.0f, axis == 2 ? 0.001f : 0.0f);
        const float v = sdf_terrain_map_detail(p + d) - sdf_terrain_map_detail(p - d);
        if (axis == 0) f[0] = v; else if (axis == 1) f[1] = v; else f[2] = v;
    }
    return normalize(vec3(f[0], f[1], f[2]));
}

// ---- lighting ----------------------------------------------------------------------------------
SBX_FN vec3 setup_lights(_in(vec3) L, _in(vec3) normal) {   // :219-238
    vec3 diffuse = vec3(0.0f, 0.0f, 0.0f);
    const vec3 c_L = vec3(7.0f, 5.0f, 3.0f);                 // key light
    diffuse += max(0.0f, dot(L, normal)) * c_L;
    const float hemi = clamp(.25f + .5f * normal.y, .0f, 1.0f);   // fill light 1 - faked hemisphere
    diffuse += hemi * vec3(.4f, .6f, .8f) * .2f;
    const float amb = clamp(.12f + .8f * max(0.0f, dot(-L, normal)), 0.0f, 1.0f);   // fill light 2 - ambient
    diffuse += amb * vec3(.4f, .5f, .6f);
    return diffuse;
}

SBX_FN vec3 illuminate(_in(vec3) pos, _in(mat3) local_xform, _in(vec2) df) {   // :240-302
    const float h = df.y;
    const vec3 w_normal = normalize(pos);
    const vec3 normal = sdf_terrain_normal(pos);
    const float N = dot(normal, w_normal);

    const vec3 c_water = vec3(.015f, .110f, .455f), c_grass = vec3(.086f, .132f, .018f), c_beach = vec3(.153f, .172f, .121f),
               c_rock = vec3(.080f, .050f, .030f), c_snow = vec3(.600f, .600f, .600f);
    const float l_water = .05f, l_shore = .17f, l_grass = .211f, l_rock = .351f;

    const float s = smoothstep(.4f, 1.0f, h);
    const vec3 rock = mix(c_rock, c_snow, smoothstep(1.0f - .3f * s, 1.0f - .2f * s, N));
    const vec3 grass = mix(c_grass, rock, smoothstep(l_grass, l_rock, h));
    vec3 shoreline = mix(c_beach, grass, smoothstep(l_shore, l_grass, h));
    const vec3 water = mix(c_water / 2.0f, c_water, smoothstep(0.0f, l_water, h));

    const vec3 L = mul(local_xform, normalize(vec3(1.0f, 1.0f, 0.0f)));
    shoreline *= setup_lights(L, normal);
    const vec3 ocean = setup_lights(L, w_normal) * water;
    return mix(ocean, shoreline, smoothstep(l_water, l_shore, h));
}

// ---- rendering ---------------------------------------------------------------------------------
SBX_FN vec3 render(_in(ray_t) eye, _in(vec3) point_cam) {   // :307-367
    sphere_t atmosphere;
    atmosphere.origin = vec3(0.0f, 0.0f, 0.0f);
    atmosphere.radius = 1.0f;
    atmosphere.material = 0;
    atmosphere.radius += max_height;

    hit_t hit = no_hit;
    intersect_sphere(eye, atmosphere, hit);
    if (hit.material_id < 0) return background(eye);

    const mat3 rot_y = rotate_around_y(27.0f);
    const mat3 rot = mul(rotate_around_x(u_time * -12.0f), rot_y);
    const mat3 rot_cloud = mul(rotate_around_x(u_time * 8.0f), rot_y);

    float t = 0.0f;
    vec2 df = vec2(1.0f, max_height);
    vec3 pos;
    float max_cld_ray_dist = max_ray_dist;
    for (int i = 0; i < TERR_STEPS; i++) {
        if (t > max_ray_dist) break;
        const vec3 o = hit.origin + t * eye.direction;
        pos = mul(rot, o);
        df = sdf_terrain_map(pos);
        if (df.x < TERR_EPS) {
            max_cld_ray_dist = t;
            break;
        }
        t += df.x * .4567f;
    }

    cloud = construct_volume(hit.origin);
    clouds_march(eye, max_cld_ray_dist, rot_cloud);

    if (df.x < TERR_EPS) {
        const vec3 c_terr = illuminate(pos, rot, df);
        const vec3 c_cld = cloud.radiance;
        const float alpha = cloud.alpha;
        // clouds ground shadows
        pos = mul(transpose(rot), pos);
        cloud = construct_volume(pos);
        const vec3 local_up = normalize(pos);
        clouds_shadow_march(local_up, rot_cloud);
        const float shadow = mix(.7f, 1.0f, step(cloud.alpha, 0.33f));
        return abs(mix(c_terr * shadow, c_cld, alpha));
    }
    return abs(mix(background(eye), cloud.radiance, cloud.alpha));
}

#define FOV tan(radians(30.0f))   // :369
#include "main.h"
