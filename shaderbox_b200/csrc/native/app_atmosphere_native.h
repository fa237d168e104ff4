// app_atmosphere_native.h -- hand-written sm_100a version of APP_ATMOSPHERE (src/app_atmosphere.h), BASELINE.json's
// 1920x1080 Rayleigh/Mie configuration.  Same plugin contract as a shaderbox app header, written against the device
// operator library; built by nvcc into images/APP_ATMOSPHERE.native.cubin.  The frame is BIT-IDENTICAL to the
// unchanged reference header compiled as a plugin: every value is produced by the reference's operations in the
// reference's order.  What is hand-tuned is how many instructions those operations take -- ncu on the plugin
// (profiles/r02a_ncu_full_atmosphere1080_plugin.json): issue slots 80.6 % busy, 9 888 warp instructions per pixel, of
// which the ~172 expf are only a third; a fifth is BSSY/BSYNC/BRA scaffolding around the slow paths of 154 IEEE
// divisions, 97 square roots and the expf range tests:
//
//  * `-height / hR` and `-height / hM` (:68-69, :125-126) divide by the constants 7994 and 1200.  With r = RN(1/d),
//    q0 = x*r, e = fma(-d, q0, x), q = fma(e, r, q0) is the correctly rounded x/d for every float with
//    2^-100 <= |x| <= 2^100 -- verified EXHAUSTIVELY against x/d for both constants (all 2^32 bit patterns; the recipe is
//    in tests/native/div_const.c, run by the CPU suite on a slice and in full by tools/).  Here x = -height with
//    height = length(sample) - earth_radius: a difference of two floats of magnitude ~6.4e6 is 0 or at least 0.25 in
//    magnitude (Sterbenz), and at most ~1.3e7, so x is inside the verified range or -0/+0.  For x = -0 the sequence
//    gives +0 where the division gives -0; the only consumer is expf, and expf(+0) == expf(-0) == 1 exactly.
//    Three FMA-pipe instructions instead of a reciprocal, four FMAs, a range check and a slow-path branch.
//  * the light march (get_sun_light, :50-76) returns before its exponentials whenever height < 0, and its samples lie
//    inside the atmosphere shell, so their arguments are in [-60000/1200 - eps, 0]: far inside expf's main range
//    (|x| < 88).  They call expf's main path directly (sbx_expf_core); the view samples (which do go underground, where
//    exp overflows) and exp(-tau) keep the full function.
//  * t1 / 16 and t1 / 8 (:59, :91) are multiplications by 0.0625 and 0.125: scaling by a power of two is exact, so
//    both forms round the same real number.
#include "def.h"
#include "util.h"
#include "intersect.h"

#define hg_g (.76f)
#include "volumetric.h"

#define earth_radius 6360e3f        // :37-38 (m)
#define atmosphere_radius 6420e3f
#define atm_sun_power 20.0f         // :41
#define atm_num_samples 16          // :47-48
#define atm_num_samples_light 8

vec3 sun_dir;                       // :40, rotated by setup_scene

// x / d for a constant d and 2^-100 <= |x| <= 2^100 (or x = +-0, see the header): r must be RN(1/d)
SBX_FN float sbx_div_const(float x, float d, float r) {
    const float q0 = x * r;
    const float e = __fmaf_rn(-d, q0, x);
    return __fmaf_rn(e, r, q0);
}
#define SBX_DIV_HR(x) sbx_div_const((x), 7994.0f, 0x1.06573cp-13f)   // hR, :33
#define SBX_DIV_HM(x) sbx_div_const((x), 1200.0f, 0x1.b4e81cp-11f)   // hM, :34

// isect_sphere (:15-26) against the atmosphere shell centred on the origin; only t1 and the hit test are used
SBX_FN bool isect_atmosphere(_in(vec3) origin, _in(vec3) direction, float& t1) {
    const vec3 rc = vec3(0.0f, 0.0f, 0.0f) - origin;
    const float radius2 = atmosphere_radius * atmosphere_radius;
    const float tca = dot(rc, direction);
    const float d2 = dot(rc, rc) - tca * tca;
    const float thc = sqrt(radius2 - d2);
    t1 = tca + thc;
    return d2 < radius2;
}

// get_sun_light (:50-76)
SBX_FN bool get_sun_light(_in(vec3) origin, float& optical_depthR, float& optical_depthM) {
    float t1;
    isect_atmosphere(origin, sun_dir, t1);
    float march_pos = 0.0f;
    const float march_step = t1 * 0.125f;                         // t1 / float(num_samples_light)
#pragma unroll 1
    for (int i = 0; i < atm_num_samples_light; i++) {
        const vec3 s = origin + sun_dir * (march_pos + 0.5f * march_step);
        const float height = length(s) - earth_radius;
        if (height < 0.0f) return false;
        optical_depthR += sbx_expf_core(SBX_DIV_HR(-height)) * march_step;
        optical_depthM += sbx_expf_core(SBX_DIV_HM(-height)) * march_step;
        march_pos += march_step;
    }
    return true;
}

// get_incident_light (:78-160)
SBX_FN vec3 get_incident_light(_in(ray_t) ray) {
    const vec3 betaR = vec3(5.5e-6f, 13.0e-6f, 22.4e-6f);         // :29-30
    const vec3 betaM = vec3(21e-6f, 21e-6f, 21e-6f);
    float t1;
    if (!isect_atmosphere(ray.origin, ray.direction, t1)) return vec3(0.0f, 0.0f, 0.0f);
    const float march_step = t1 * 0.0625f;                        // t1 / float(num_samples)
    const float mu = dot(ray.direction, sun_dir);
    const float phaseR = rayleigh_phase_func(mu);
    const float phaseM = henyey_greenstein_phase_func(mu);
    float optical_depthR = 0.0f, optical_depthM = 0.0f;
    vec3 sumR = vec3(0.0f, 0.0f, 0.0f), sumM = vec3(0.0f, 0.0f, 0.0f);
    float march_pos = 0.0f;
#pragma unroll 1
    for (int i = 0; i < atm_num_samples; i++) {
        const vec3 s = ray.origin + ray.direction * (march_pos + 0.5f * march_step);
        const float height = length(s) - earth_radius;
        // integrate the height scale (these samples do go below ground: the full exp)
        const float hr = exp(SBX_DIV_HR(-height)) * march_step;
        const float hm = exp(SBX_DIV_HM(-height)) * march_step;
        optical_depthR += hr;
        optical_depthM += hm;
        // gather the sunlight
        float optical_depth_lightR = 0.0f, optical_depth_lightM = 0.0f;
        if (get_sun_light(s, optical_depth_lightR, optical_depth_lightM)) {
            const vec3 tau = betaR * (optical_depthR + optical_depth_lightR) + betaM * 1.1f * (optical_depthM + optical_depth_lightM);
            const vec3 attenuation = exp(-tau);
            sumR += hr * attenuation;
            sumM += hm * attenuation;
        }
        march_pos += march_step;
    }
    return atm_sun_power * (sumR * phaseR * betaR + sumM * phaseM * betaM);
}

SBX_FN void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) {   // :164-175 (FROM_SPACE)
    eye = vec3(0.0f, 0.0f, 0.0f);
    look_at = vec3(0.0f, 1.0f, 0.0f);
}

SBX_FN void setup_scene() {                                          // :177-181
    const mat3 rot = rotate_around_x(-abs(sin(u_time / 2.0f)) * 90.0f);
    sun_dir = mul(vec3(0.0f, 1.0f, 0.0f), rot);
}

SBX_FN vec3 render(_in(ray_t) eye, _in(vec3) point_cam) {            // :183-228 (FROM_SPACE): sky-dome angles
    const vec3 p = point_cam;
    const float z2 = p.x * p.x + p.y * p.y;
    const float phi = atan(p.y, p.x);
    const float theta = acos(1.0f - z2);
    const vec3 dir = vec3(sin(theta) * cos(phi), cos(theta), sin(theta) * sin(phi));
    ray_t ray;
    ray.origin = vec3(0.0f, earth_radius + 1.0f, 0.0f);
    ray.direction = dir;
    return get_incident_light(ray);
}

#define FOV 1.0f   // :230
#include "main.h"
