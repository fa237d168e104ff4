// app_clouds_native.h -- hand-written sm_100a version of APP_CLOUDS (src/app_clouds.h), the app the
// headline metric is quoted on.  Same plugin contract as a shaderbox app header (setup_camera /
// setup_scene / render / FOV, then main.h), written directly against the device operator library;
// built by nvcc into images/APP_CLOUDS.native.cubin.  The frame is BIT-IDENTICAL to the unchanged
// reference header compiled as a plugin (tests/test_gpu_parity.py compares both with the oracle):
// every value that reaches the pixel is produced by the reference's operations in the reference's
// order.  What is hand-tuned is which of those operations are executed at all, and how many issue
// slots the remaining ones take (ncu, profiles/: the kernel is instruction-issue bound, ~90 % of
// issued instructions are the fbm octaves of density_func):
//
//  * packed fp32 (FFMA2).  The octave arithmetic runs two lanes per instruction: x/y for the
//    position chain and the smoothstep weights, the two z-slices of a noise cell for the x/y
//    interpolation (the memo table stores z-neighbours adjacent, noise_iq.h), octave pairs in the
//    light march.  Every lane performs exactly the scalar operation's single rounding
//    (sbx_vec.cuh, pk_*).
//  * lazy octaves in the view march (density_func :62-86).  density = shape * smoothstep(cov,
//    cov + .0135, shape) is exactly +0 whenever shape <= cov, and the fbm octaves still to come can
//    add at most their gains (noise_iq <= 1): after octave i, t_i + sum(remaining H) <= cov (minus a
//    1e-5 guard, orders of magnitude above the few ulps the fp32 sums can gain) proves the result
//    is 0 without evaluating them.  Half of the sky is empty (cld_coverage .535).
//  * smoothstep without its division outside the 0.0135-wide band: with a = shape - cov and
//    b = (cov + .0135) - cov > 0 (both as rounded by the reference), a <= 0 gives t = 0 and the
//    density +0; a >= b gives a/b >= 1 (division is monotonic, 1 is representable), t = 1 and
//    the density shape * (1*1*(3-2)) = shape.  Only samples inside the band divide.
//  * z-slice reuse in the light march (illuminate_volume :107-113).  noise_iq interpolates x, then
//    y, then z (src/noise_iq.h:19-23): the two bilinear z-slice values y0, y1 of an octave depend
//    only on (p.x, p.y, floor(p.z)).  The light march steps by L*dt; when that leaves x and y
//    bit-unchanged (the default sun_dir (0,0,-1) of src/uniform_buffer.h:42) an octave whose z
//    stays in its lattice cell is one floor, one smoothstep weight and one mix; an octave that
//    crossed into the next cell is re-sliced alone.  Equality is checked on the values, per sample
//    and per octave, so any sun direction stays exact; it just re-slices every octave.
//  * the Henyey-Greenstein factor of illuminate_volume (:121) depends only on the ray, not on the
//    sample: one pow per ray instead of one per in-cloud step.
//  * table misses are deferred: lattice indices are clamped into the memo table and the largest
//    one is tracked; only a pixel that ever went out of range is recomputed, on the generic path
//    (no calls and no miss branches inside the march loops).
#include "def.h"
#include "util.h"
#include "intersect.h"

// Issue-order hint for the launcher (sbx_kernel.cuh, SBX_HINT_TRIVIAL_ROWS is set on the compiler command line from this
// value, csrc/Makefile): with the camera of setup_camera (:23-30: eye (0,-.5,0) looking at (0,0,-1), yaw only) and FOV 1,
// a ray's direction has y < 0.05 -- render() returns the sky colour, :212 -- on every pixel of the bottom 25.1 % of the
// frame (centre column; 28 % at the left and right edges).  250 per mille of the rows are issued last.  A performance
// hint only: pixels there that do meet cloud (another camera) are marched like any other.

#define hg_g (.2f)
#include "volumetric.h"
#include "noise_iq.h"
#include "fbm.h"

#define cld_noise_factor .001f
DECL_FBM_FUNC(fbm, 4, noise_iq(p))   // :59, the generic form: used when a lattice index leaves the memo table

float sbx_phase;       // henyey_greenstein_phase_func(clamp(dot(L, V), 0, 1)) of :121, per ray
float sbx_cov;         // 1 - cld_coverage (:83)
float sbx_band_lo, sbx_band_hi;   // smoothstep shortcuts: (shape - cov) <= lo -> 0, >= hi -> 1 (see sbx_density_of)
// z-slice memo of the last sliced sample: position x/y it was taken at and, per octave, the
// lattice z and the two bilinear slice values (y0, y1)
float sbx_mx, sbx_my;
float sbx_mz0, sbx_mz1, sbx_mz2, sbx_mz3;
float2 sbx_ys0, sbx_ys1, sbx_ys2, sbx_ys3;
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

// The two z-slices (y0, y1) of one noise_iq octave (src/noise_iq.h:11-23) at lattice cell
// (px, py, pz) with smoothed x/y weights.  The lattice index is clamped into the memo table and
// the largest one remembered: a pixel that ever asked for an entry outside the table is
// recomputed on the generic path (render()).
SBX_FN float2 sbx_slices(float2 pxy, float pz, float2 wxy) {
    const float n = pxy.x + pxy.y * 157.0f + 113.0f * pz;
    const unsigned k = (unsigned)(__float_as_int(n + 12582912.0f) - sbx_L->hash_bias);
    sbx_kmax = ::max(sbx_kmax, k);
    const float4* __restrict__ e = sbx_L->hash_tab + 2u * ::min(k, (unsigned)sbx_L->hash_span - 1u);
    const float4 lo = __ldg(e), hi = __ldg(e + 1);
    const float2 axy = pk_one_minus(wxy);
    return pk_mix(pk_mix(pk(lo.x, lo.y), pk(lo.z, lo.w), axy.x, wxy.x),
                  pk_mix(pk(hi.x, hi.y), pk(hi.z, hi.w), axy.x, wxy.x), axy.y, wxy.y);
}

// -DSBX_X_FLOOR_MAGIC (experiment, profiles/r02g): floor() without the XU pipe's FRND -- (x + 1.5*2^23) - 1.5*2^23 is x
// rounded to the nearest integer for |x| < 2^22 (larger lattice coordinates leave the memo table anyway and take the
// generic path), minus 1 where that rounded up.  Same value as floorf except the sign of a zero result.
#ifdef SBX_X_FLOOR_MAGIC
SBX_FN float sbx_floor(float x) {
    const float n = (x + 12582912.0f) - 12582912.0f;
    return n > x ? n - 1.0f : n;
}
#else
SBX_FN float sbx_floor(float x) { return floor(x); }
#endif

// one octave at p = (pxy, pz), evaluated in full; leaves its lattice z and slices in the memo
SBX_FN float sbx_octave(float2 pxy, float pz, float& mz, float2& ys) {
    const float2 cxy = pk(sbx_floor(pxy.x), sbx_floor(pxy.y));
    const float cz = sbx_floor(pz);
    ys = sbx_slices(cxy, cz, sbx_noise_weight(pk_sub(pxy, cxy)));
    mz = cz;
    return sbx_noise_zmix(ys, sbx_noise_weight(pz - cz));
}

// shape * smoothstep(cov, cov + .0135, shape)   (:83-84), dividing only inside the band
SBX_FN float sbx_density_of(float shape) {
    const float a = shape - sbx_cov;
    if (a <= sbx_band_lo) return 0.0f;
    if (a >= sbx_band_hi) return shape;
    return shape * smoothstep(sbx_cov, sbx_cov + .0135f, shape);
}

// density_func (:62-86) for a view-march sample: fbm(pos * 2.03, 2.64, .5, .5) of src/fbm.h:6
// unrolled (H = .5 .25 .125 .0625) with the lazy-octave exits
SBX_FN float density_view(float2 pos_xy, float pos_z) {
    float2 pxy = pk_mul(pk_mul(pos_xy, cld_noise_factor), 2.03f);     // pos = pos_in * .001; p = pos * 2.03
    float pz = pos_z * cld_noise_factor * 2.03f;
    const float guard = 1e-5f;
    sbx_mx = pos_xy.x; sbx_my = pos_xy.y;
    float t = sbx_octave(pxy, pz, sbx_mz0, sbx_ys0) * .5f;            // 0 + n*.5 == n*.5 (n >= +0)
#if !defined(SBX_X_NOLAZY) && defined(SBX_X_LAZY1)
    // exit after the first octave: provable, but it fires on 2 % of the samples (ncu source counters, profiles/r02c) and
    // keeps the second octave's table loads from being issued beside the first's: without it +1 % (1 lane), +2 % (4 lanes)
    if (t <= sbx_cov - .4375f - guard) return 0.0f;
#endif
    pxy = pk_mul(pxy, 2.64f); pz *= 2.64f;
    t += sbx_octave(pxy, pz, sbx_mz1, sbx_ys1) * .25f;
#ifndef SBX_X_NOLAZY
    if (t <= sbx_cov - .1875f - guard) return 0.0f;
#endif
    pxy = pk_mul(pxy, 2.64f); pz *= 2.64f;
    t += sbx_octave(pxy, pz, sbx_mz2, sbx_ys2) * .125f;
#ifndef SBX_X_NOLAZY
    if (t <= sbx_cov - .0625f - guard) return 0.0f;
#endif
    pxy = pk_mul(pxy, 2.64f); pz *= 2.64f;
    t += sbx_octave(pxy, pz, sbx_mz3, sbx_ys3) * .0625f;
    return sbx_density_of(t);
}

// exp(-density * sigma * dt) of :110 and :135 (Beer-Lambert).  density lies in [0, 1) (noise values times gains
// that add up to .9375), so |argument| <= |sigma * dt|: when that is below 80 -- decided once per ray, render()
// picks the BOUNDED instantiation of the march -- expf's range test cannot fire and its main path is called directly.
template <bool BOUNDED> SBX_FN float sbx_beer_lambert(float density, float dt) {
    const float x = -density * sigma_scattering * dt;
    return BOUNDED ? sbx_expf_core(x) : exp(x);
}

// illuminate_volume (:91-123).  `origin` is the view sample just evaluated by density_view, so
// the memo holds its slices.  Per light sample: the z chain of the four octaves, their position in the
// memoised lattice cells, and -- if x and y still equal the memo's -- only the octaves that left their
// cell are re-sliced.  Then every octave is one weight (two octaves per FFMA2) and one mix of its two slices.
template <bool BOUNDED> SBX_FN float illuminate_volume(float2 origin_xy, float origin_z, _in(vec3) L) {
    const float dt = cld_thick / float(cld_march_steps);
    const vec3 step = L * dt;
    float2 pos_xy = pk_add(origin_xy, pk(step.x, step.y));     // don't sample just where the main raymarcher is
    float pos_z = origin_z + step.z;
    float transmittance = 1.0f;
    float2 ys0 = sbx_ys0, ys1 = sbx_ys1, ys2 = sbx_ys2, ys3 = sbx_ys3;
    float2 m01 = pk(sbx_mz0, sbx_mz1), m23 = pk(sbx_mz2, sbx_mz3);
    for (int i = 0; i < illum_march_steps; i++) {
        const float z0 = pos_z * cld_noise_factor * 2.03f;     // the z chain of p = pos*.001*2.03, p *= 2.64
        const float z1 = z0 * 2.64f, z2 = z1 * 2.64f, z3 = z2 * 2.64f;
        const float2 z01 = pk(z0, z1), z23 = pk(z2, z3);
        // fract(z) against the memoised lattice z: f = z - m lies in [0, 1) only if sbx_floor(z) == m (z < m gives
        // f < 0, z >= m + 1 gives f >= 1: rounding is monotonic), and then it IS the reference's z - sbx_floor(z).
        // As unsigned integers the floats of [+0, 1) are exactly the values below bits(1.0f).
        float2 f01 = pk_sub(z01, m01), f23 = pk_sub(z23, m23);
        const unsigned worst = ::max(::max(__float_as_uint(f01.x), __float_as_uint(f01.y)),
                                     ::max(__float_as_uint(f23.x), __float_as_uint(f23.y)));
        const bool same_xy = pos_xy.x == sbx_mx && pos_xy.y == sbx_my;
        if (!(same_xy && worst < 0x3f800000u)) {
            // re-slice the octaves that left their cell (all of them if x or y changed)
            const float2 c01 = pk(sbx_floor(z0), sbx_floor(z1)), c23 = pk(sbx_floor(z2), sbx_floor(z3));
            float2 pxy = pk_mul(pk_mul(pos_xy, cld_noise_factor), 2.03f);
            float2 cxy;
            if (!(same_xy && c01.x == m01.x)) {
                cxy = pk(sbx_floor(pxy.x), sbx_floor(pxy.y));
                ys0 = sbx_slices(cxy, c01.x, sbx_noise_weight(pk_sub(pxy, cxy)));
                m01.x = c01.x;
            }
            pxy = pk_mul(pxy, 2.64f);
            if (!(same_xy && c01.y == m01.y)) {
                cxy = pk(sbx_floor(pxy.x), sbx_floor(pxy.y));
                ys1 = sbx_slices(cxy, c01.y, sbx_noise_weight(pk_sub(pxy, cxy)));
                m01.y = c01.y;
            }
            pxy = pk_mul(pxy, 2.64f);
            if (!(same_xy && c23.x == m23.x)) {
                cxy = pk(sbx_floor(pxy.x), sbx_floor(pxy.y));
                ys2 = sbx_slices(cxy, c23.x, sbx_noise_weight(pk_sub(pxy, cxy)));
                m23.x = c23.x;
            }
            pxy = pk_mul(pxy, 2.64f);
            if (!(same_xy && c23.y == m23.y)) {
                cxy = pk(sbx_floor(pxy.x), sbx_floor(pxy.y));
                ys3 = sbx_slices(cxy, c23.y, sbx_noise_weight(pk_sub(pxy, cxy)));
                m23.y = c23.y;
            }
            sbx_mx = pos_xy.x; sbx_my = pos_xy.y;
            f01 = pk_sub(z01, m01); f23 = pk_sub(z23, m23);
        }
        const float2 w01 = sbx_noise_weight(f01), w23 = sbx_noise_weight(f23);
        const float2 a01 = pk_one_minus(w01), a23 = pk_one_minus(w23);
        // noise_iq of each octave = mix(y0, y1, w) on its two memoised slices, then the fbm sum (src/fbm.h:6)
        const float n0 = ys0.x * a01.x + ys0.y * w01.x, n1 = ys1.x * a01.y + ys1.y * w01.y;
        const float n2 = ys2.x * a23.x + ys2.y * w23.x, n3 = ys3.x * a23.y + ys3.y * w23.y;
        const float density = sbx_density_of(((n0 * .5f + n1 * .25f) + n2 * .125f) + n3 * .0625f);
#ifdef SBX_X_NOEXP
        if (density != 0.0f) transmittance *= (1.0f - density * sigma_scattering * dt);
#else
        transmittance *= sbx_beer_lambert<BOUNDED>(density, dt);   // branch-free: an empty sample multiplies by exp(-0) == 1
#endif
        pos_xy = pk_add(pos_xy, pk(step.x, step.y));
        pos_z += step.z;
    }
    return transmittance * sun_power * sbx_phase;
}

template <bool BOUNDED> SBX_FN vec4 render_clouds(_in(ray_t) eye) {   // :153-202
    const vec3 projection = eye.direction / eye.direction.y;
    vec3 origin = eye.origin + projection * 150.0f;
    origin += wind_dir * u_time * (1.0f / cld_noise_factor);

    sbx_cov = 1.0f - cld_coverage;
    const float band = (sbx_cov + .0135f) - sbx_cov;                      // e1 - e0 of the smoothstep (:84)
    sbx_band_lo = band > 0.0f ? 0.0f : -__int_as_float(0x7f800000);        // a degenerate band never shortcuts
    sbx_band_hi = band > 0.0f ? band : __int_as_float(0x7f800000);
    sbx_phase = henyey_greenstein_phase_func(clamp(dot(sun_dir, eye.direction), 0.0f, 1.0f));

    const float2 origin_xy = pk(origin.x, origin.y), proj_xy = pk(projection.x, projection.y);
    float transmittance = 1.0f, radiance = 0.0f, alpha = 0.0f;             // volume_sampler_t (volumetric.h:47-68);
    float t = 0.0f;                                                        // radiance is one float: vec3 += float
    const float dt = cld_thick / float(cld_march_steps);
    for (int i = 0; i < cld_march_steps; i++) {
        const float2 pos_xy = pk_add(origin_xy, pk_mul(proj_xy, t));       // cloud.pos = cloud.origin + t * projection
        const float pos_z = origin.z + t * projection.z;
        t += dt;
        const float density = density_view(pos_xy, pos_z);
        if (!(density < .005f)) {                                          // integrate_volume, :125-148
            const float T_i = sbx_beer_lambert<BOUNDED>(density, dt);      // Beer-Lambert
            transmittance *= T_i;
#ifdef SBX_X_NOLIGHT
            radiance += (density * sigma_scattering) * (sbx_ys0.x + sbx_ys1.y + sbx_ys2.x + sbx_ys3.y + sbx_mz0 + sbx_mz1 + sbx_mz2 + sbx_mz3) * transmittance * dt;
#else
            radiance += (density * sigma_scattering) * illuminate_volume<BOUNDED>(pos_xy, pos_z, sun_dir) * transmittance * dt;
#endif
            alpha += (1.0f - T_i) * (1.0f - alpha);
        }
        if (alpha > .999f) break;
    }
    const float cutoff = dot(eye.direction, vec3(0.0f, 1.0f, 0.0f));
    return vec4(radiance, radiance, radiance, alpha * smoothstep(.0f, .2f, cutoff));
}

// lanes per pixel of the cooperative march compiled into this image: every warp (-DSBX_LANES_PER_PIXEL=P), or the
// warps of the second region of a hybrid image (-DSBX_HYBRID_LANES=P, sbx_kernel.cuh); 1 = none
#if SBX_LANES_PER_PIXEL > 1
#define SBX_COOP_LANES SBX_LANES_PER_PIXEL
#elif SBX_HYBRID_LANES > 1
#define SBX_COOP_LANES SBX_HYBRID_LANES
#else
#define SBX_COOP_LANES 1
#endif
#if SBX_COOP_LANES > 1
// Cooperative march (image built with -DSBX_LANES_PER_PIXEL=P or -DSBX_HYBRID_LANES=P): the P lanes of a pixel take the
// view-march steps i = r*P + phase of round r.  Everything expensive in a step -- its density,
// its Beer-Lambert factor T_i and its light march -- depends only on the step's position, not on
// the accumulators, so the lanes evaluate P steps of the SAME ray at once; the accumulators
// (transmittance, radiance, alpha; :138-147) are then updated in step order from warp shuffles,
// redundantly in all P lanes, with the reference's operations.  Steps evaluated past the
// alpha > .999 exit (:197) are discarded.  A ray's work is spread over P times as many warps, each
// P times shorter: frames that are small for the machine (one GPU's share of a 1080p frame at 8
// GPUs is ~2 waves of warps) no longer end in a long single-warp tail, and the lanes of a warp
// sit on neighbouring steps of the same rays, which keeps their branches and table lines together.
// The three pows a cloud pixel needs before its march -- the two sun lobes of render_sky_color (:42-43) and the
// Henyey-Greenstein denominator (volumetric.h:27-33 via :121) -- are independent: the lanes of a pixel compute one
// each (same function, same arguments as the one-lane path, so the same bits) and pass them round with shuffles.
SBX_FN void sbx_coop_pows(_in(vec3) eye_dir, float& lobe1500, float& lobe10, float& hg_pow) {
    const int P = SBX_COOP_LANES;
    const int lane = threadIdx.x & 31, phase = lane % P, base = lane - phase;
    const float sun_amount = max(dot(eye_dir, sun_dir), 0.0f);
    const float mu = clamp(dot(sun_dir, eye_dir), 0.0f, 1.0f);
    float mine[(3 + P - 1) / P];
#pragma unroll
    for (int k = 0; k * P < 3; ++k) {
        const int which = k * P + phase;            // 0: sun_amount^1500, 1: sun_amount^10, 2 (and any spare lane): the HG power
        const float b = which < 2 ? sun_amount : 1.0f + hg_g * hg_g - 2.0f * hg_g * mu;
        const float e = which == 0 ? 1500.0f : which == 1 ? 10.0f : 1.5f;
        mine[k] = pow(b, e);
    }
    lobe1500 = __shfl_sync(0xffffffffu, mine[0], base);
    lobe10 = __shfl_sync(0xffffffffu, mine[1 / P], base + 1 % P);
    hg_pow = __shfl_sync(0xffffffffu, mine[2 / P], base + 2 % P);
}

// linear_to_srgb (src/util.h:72-77 via src/main.h:52) for a pixel marched by several lanes: one channel per lane
SBX_FN vec3 sbx_coop_encode(_in(vec3) color) {
    const int P = SBX_COOP_LANES;
    const int lane = threadIdx.x & 31, phase = lane % P, base = lane - phase;
    float mine[(3 + P - 1) / P];
#pragma unroll
    for (int k = 0; k * P < 3; ++k) {
        const int which = k * P + phase;
        const float c = which == 0 ? color.x : which == 1 ? color.y : color.z;
        mine[k] = pow(c, 1.0f / 2.2f);
    }
    return vec3(__shfl_sync(0xffffffffu, mine[0], base), __shfl_sync(0xffffffffu, mine[1 / P], base + 1 % P),
                __shfl_sync(0xffffffffu, mine[2 / P], base + 2 % P));
}

template <bool BOUNDED> SBX_FN vec4 render_clouds_coop(_in(ray_t) eye, bool active, float hg_pow) {
    const int P = SBX_COOP_LANES;
    const int lane = threadIdx.x & 31, phase = lane % P, base = lane - phase;
    const vec3 projection = eye.direction / eye.direction.y;
    vec3 origin = eye.origin + projection * 150.0f;
    origin += wind_dir * u_time * (1.0f / cld_noise_factor);

    sbx_cov = 1.0f - cld_coverage;
    const float band = (sbx_cov + .0135f) - sbx_cov;
    sbx_band_lo = band > 0.0f ? 0.0f : -__int_as_float(0x7f800000);
    sbx_band_hi = band > 0.0f ? band : __int_as_float(0x7f800000);
    sbx_phase = (1.0f - hg_g * hg_g) / ((4.0f + PI) * hg_pow);   // henyey_greenstein_phase_func (volumetric.h:27-33) on the shared pow

    const float2 origin_xy = pk(origin.x, origin.y), proj_xy = pk(projection.x, projection.y);
    float transmittance = 1.0f, radiance = 0.0f, alpha = 0.0f;
    const float dt = cld_thick / float(cld_march_steps);
    const int steps = cld_march_steps;
    float t = 0.0f;                                    // the reference's t of step `phase`: 0 (+ dt) (+ dt) ...
    for (int q = 0; q < phase; ++q) t += dt;
    bool done = !active;
    // A round = U sub-rounds of P consecutive steps (lane `phase` takes step (r*U + h)*P + phase in sub-round h), evaluated
    // back to back, then ONE vote and one in-order merge of the round's U*P steps; steps evaluated past the alpha exit are
    // discarded (at most U*P - 1 of them per ray).  U = 2 halves the votes per step but measured 5 % SLOWER than U = 1
    // (one rank's share at 8 GPUs 0.358 vs 0.342 ms, full frame 2.635 vs 2.504 ms; profiles/r02o_*): U stays 1.
#ifndef SBX_COOP_UNROLL
#define SBX_COOP_UNROLL 1
#endif
    const int U = SBX_COOP_UNROLL;
    for (int r = __all_sync(0xffffffffu, done) ? steps : 0; r * P * U < steps; ++r) {   // (a warp of rays below the horizon marches nothing)
        float T_h[U], A_h[U];
        unsigned hits_h[U];
        bool any_hit = false;
#pragma unroll
        for (int h = 0; h < U; ++h) { T_h[h] = 1.0f; A_h[h] = 0.0f; }
#pragma unroll 1
        for (int h = 0; h < U; ++h) {
            const int i = (r * U + h) * P + phase;
            float T_i = 1.0f, A = 0.0f;
            bool hit = false;                          // this lane's step met cloud (integrate_volume's density test, :132)
            if (!done && i < steps) {
                const float2 pos_xy = pk_add(origin_xy, pk_mul(proj_xy, t));
                const float pos_z = origin.z + t * projection.z;
                const float density = density_view(pos_xy, pos_z);
                if (!(density < .005f)) {
                    hit = true;
                    T_i = sbx_beer_lambert<BOUNDED>(density, dt);
                    A = (density * sigma_scattering) * illuminate_volume<BOUNDED>(pos_xy, pos_z, sun_dir);
                }
            }
#pragma unroll
            for (int q = 0; q < P; ++q) t += dt;       // this lane's next step is P steps on
            any_hit = any_hit || hit;
            const unsigned hits = hit ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < U; ++k)                // (a select per slot keeps T_h / A_h in registers under the rolled loop)
                if (k == h) { T_h[k] = T_i; A_h[k] = A; hits_h[k] = hits; }
        }
        if (__any_sync(0xffffffffu, any_hit)) {        // rounds in which no lane of the warp met cloud change nothing
#pragma unroll
            for (int h = 0; h < U; ++h) {
                const unsigned hits = __ballot_sync(0xffffffffu, hits_h[h] != 0u) >> base;   // bit q: step (r*U+h)*P + q integrates
#pragma unroll
                for (int q = 0; q < P; ++q) {          // the round's steps in order
                    const float Tq = __shfl_sync(0xffffffffu, T_h[h], base + q);
                    const float Aq = __shfl_sync(0xffffffffu, A_h[h], base + q);
                    // branch-free: a step that integrates nothing multiplies by 1 and adds +0, which change no bit of
                    // the accumulators whatever they hold (x * 1 == x; x + 0 == x for every x but -0, and none of them can
                    // be -0: they start at 1, +0, +0 and a sum only yields -0 from two -0 operands)
                    const bool use = !done && ((hits >> q) & 1u);
                    transmittance *= use ? Tq : 1.0f;
                    radiance += use ? Aq * transmittance * dt : 0.0f;
                    alpha += use ? (1.0f - Tq) * (1.0f - alpha) : 0.0f;
                    done = done || (use && alpha > .999f); // :197 (alpha only changes in a cloud step)
                }
            }
            if (__all_sync(0xffffffffu, done)) break;  // `done` only changes in a merge: the exit vote is taken here alone
        }
    }
#pragma unroll
    for (int off = P / 2; off > 0; off >>= 1) sbx_kmax = ::max(sbx_kmax, __shfl_xor_sync(0xffffffffu, sbx_kmax, off));
    const float cutoff = dot(eye.direction, vec3(0.0f, 1.0f, 0.0f));
    return vec4(radiance, radiance, radiance, alpha * smoothstep(.0f, .2f, cutoff));
}
#endif

// ---- the generic path: the app as written (:62-202) on the library's noise_iq / fbm, which fall
// back to the arithmetic hash outside the memo table.  Cold code: runs only for a pixel whose fast
// path met a lattice index outside the table (huge u_time * wind_dir, or the table switched off).
SBX_FN float generic_density(_in(vec3) pos_in) {
    const vec3 pos = pos_in * cld_noise_factor;
    const float shape = fbm(pos * 2.03f, 2.64f, .5f, .5f);
    return shape * smoothstep(sbx_cov, sbx_cov + .0135f, shape);
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
    sbx_kmax = 0u;
    const bool bounded = abs(sigma_scattering * (cld_thick / float(cld_march_steps))) < 80.0f;   // see sbx_beer_lambert
    vec3 sky;
    vec4 cld;
#if SBX_COOP_LANES > 1
    if (sbx_coop) {   // warp-uniform (a compile-time constant unless the image is a hybrid)
        float lobe1500, lobe10, hg_pow;
        sbx_coop_pows(eye_ray.direction, lobe1500, lobe10, hg_pow);
        sky = mix(vec3(.0f, .1f, .4f), vec3(.3f, .6f, .8f), 1.0f - eye_ray.direction.y);   // render_sky_color (:36-46)
        sky += sun_color * min(lobe1500 * 5.0f, 1.0f);
        sky += sun_color * min(lobe10 * .6f, 1.0f);
        sky = abs(sky);
        const bool below = dot(eye_ray.direction, vec3(0.0f, 1.0f, 0.0f)) < 0.05f;
        // (the uniforms, hence `bounded`, are the same in every lane: the whole warp takes one instantiation)
        cld = bounded ? render_clouds_coop<true>(eye_ray, !below, hg_pow) : render_clouds_coop<false>(eye_ray, !below, hg_pow);
        if (below) return sky;
    } else
#endif
    {
#if SBX_LANES_PER_PIXEL == 1
        sky = render_sky_color(eye_ray.direction);
        if (dot(eye_ray.direction, vec3(0.0f, 1.0f, 0.0f)) < 0.05f) return sky;
        cld = bounded ? render_clouds<true>(eye_ray) : render_clouds<false>(eye_ray);
#endif
    }
    if (sbx_kmax >= (unsigned)sbx_L->hash_span) {                       // table miss: redo the pixel
        const float4 g = sbx_generic_pixel(sbx_L, eye_ray.origin.x, eye_ray.origin.y, eye_ray.origin.z,
                                           eye_ray.direction.x, eye_ray.direction.y, eye_ray.direction.z);
        cld = vec4(g.x, g.y, g.z, g.w);
    }
    const vec3 col = mix(sky, cld.rgb, cld.a);
    return abs(col);
}

#define FOV 1.0f   // :220
#if SBX_COOP_LANES > 1
// main.h's sRGB encode (src/main.h:52), one channel per lane when several lanes march the pixel
#define SBX_APP_ENCODE(color) (sbx_coop ? sbx_coop_encode(color) : linear_to_srgb(color))
#endif
#include "main.h"
