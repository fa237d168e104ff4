/* sbx_oracle.c -- TEST INFRASTRUCTURE: plain-C CPU restatement of the shaderbox pixel path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
 * this file.  It is never linked into, called from or shipped with the product (libsbx.so).
 *
 * What it restates: per-pixel mainImage (src/main.h:6-53) of APP_EGG, APP_CLOUDS,
 * APP_ATMOSPHERE, APP_PLANET, APP_RAYTRACER, APP_SDF_AO and APP_VINYL together with the operator headers they use
 * (util.h, intersect.h, sdf.h, IK.h, noise_iq.h, fbm.h, volumetric.h, material.h, light.h,
 * util_optics.h, cornell_box.h).  Every function cites the reference lines it follows.
 *
 * Pinning: the reference ships no tests or golden images, so the pin is the
 * reference ITSELF run here -- oracle/_ref (the verbatim headers on oracle/ref/glsl_shim.h):
 * tests/test_oracle_cpu.py requires this file to reproduce oracle/_ref BIT FOR BIT, and
 * tests/golden/ holds frames generated from oracle/_ref by tools/make_golden.py.
 *
 * Arithmetic contract: fp32, one IEEE operation per operator in the reference's expression order,
 * no contraction, no fast-math; vector semantics per the GLSL spec formulas as listed in
 * oracle/ref/glsl_shim.h; sin cos tan exp pow acos atan2 are glibc's, called at run time
 * (see oracle/Makefile NOFOLD).  Per-pixel (GLSL) state semantics: every `_mutable` of the
 * reference starts from its initialiser at every pixel.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "../include/sbx.h"

/* ---- call counters (for the FP32 work estimate in bench.py) ---------------------------------- */
typedef struct { unsigned long long n_sin, n_cos, n_exp, n_pow, n_sqrt, n_other; } counts_t;
static _Thread_local counts_t tl;
static inline float m_sin(float a) { tl.n_sin++; return sinf(a); }
static inline float m_cos(float a) { tl.n_cos++; return cosf(a); }
static inline float m_tan(float a) { tl.n_other++; return tanf(a); }
static inline float m_exp(float a) { tl.n_exp++; return expf(a); }
static inline float m_pow(float a, float b) { tl.n_pow++; return powf(a, b); }
static inline float m_sqrt(float a) { tl.n_sqrt++; return sqrtf(a); }
static inline float m_acos(float a) { tl.n_other++; return acosf(a); }
static inline float m_atan2(float y, float x) { tl.n_other++; return atan2f(y, x); }

/* ---- the vector layer (what VML / GLSL provide), scalar form ---------------------------------- */
typedef struct { float x, y; } v2;
typedef struct { float x, y, z; } v3;
typedef struct { v3 c0, c1, c2; } m3;   /* columns */

static inline v2 V2(float x, float y) { v2 r = {x, y}; return r; }
static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v2 add2(v2 a, v2 b) { return V2(a.x + b.x, a.y + b.y); }
static inline v2 sub2(v2 a, v2 b) { return V2(a.x - b.x, a.y - b.y); }
static inline v2 scale2(v2 a, float s) { return V2(a.x * s, a.y * s); }
static inline v2 rscale2(float s, v2 a) { return V2(s * a.x, s * a.y); }
static inline v2 divs2(v2 a, float s) { return V2(a.x / s, a.y / s); }
static inline float dot2(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }     /* vec * float */
static inline v3 rscale3(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }    /* float * vec */
static inline v3 divs3(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 adds3(v3 a, float s) { return V3(a.x + s, a.y + s, a.z + s); }
static inline v3 neg3(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 abs3(v3 a) { return V3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float len2(v2 a) { return m_sqrt(dot2(a, a)); }
static inline float len3(v3 a) { return m_sqrt(dot3(a, a)); }
static inline v3 norm3(v3 a) { return divs3(a, len3(a)); }
static inline v3 cross3(v3 a, v3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float g_min(float a, float b) { return fminf(a, b); }
static inline float g_max(float a, float b) { return fmaxf(a, b); }
static inline float g_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float g_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
static inline v2 mix2(v2 x, v2 y, float a) { return V2(g_mix(x.x, y.x, a), g_mix(x.y, y.y, a)); }
static inline v3 mix3(v3 x, v3 y, float a) { return V3(g_mix(x.x, y.x, a), g_mix(x.y, y.y, a), g_mix(x.z, y.z, a)); }
static inline float g_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
static inline float g_smoothstep(float e0, float e1, float x) {
    float t = g_clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
static inline float g_fract(float a) { return a - floorf(a); }
static inline float g_radians(float d) { return d * 0.017453292519943295f; }
static inline m3 M3(float a, float b, float c, float d, float e, float f, float g, float h, float i) {
    m3 m = {{a, b, c}, {d, e, f}, {g, h, i}};
    return m;
}
static inline v3 mat_vec(m3 m, v3 v) {   /* M*v = c0*v.x + c1*v.y + c2*v.z */
    return add3(add3(scale3(m.c0, v.x), scale3(m.c1, v.y)), scale3(m.c2, v.z));
}
static inline v3 vec_mat(v3 v, m3 m) { return V3(dot3(v, m.c0), dot3(v, m.c1), dot3(v, m.c2)); }
static inline m3 mat_mat(m3 a, m3 b) {
    m3 r = {mat_vec(a, b.c0), mat_vec(a, b.c1), mat_vec(a, b.c2)};
    return r;
}
static inline m3 transpose3(m3 m) {   /* src/util.h:25-32 */
    return M3(m.c0.x, m.c1.x, m.c2.x, m.c0.y, m.c1.y, m.c2.y, m.c0.z, m.c1.z, m.c2.z);
}

/* ---- def.h ------------------------------------------------------------------------------------ */
#define PI_F 3.14159265359f      /* src/def.h:51 */
#define BIAS 1e-4f               /* src/def.h:57 */
#define MAX_DIST 1e8f            /* src/def.h:77 */
typedef struct { v3 origin, direction; } ray_t;                          /* src/def.h:53-56 */
typedef struct { v3 origin; float radius; int material; } sphere_t;      /* src/def.h:59-63 */
typedef struct { v3 direction; float distance; int material; } plane_t;  /* src/def.h:65-69 */
typedef struct { float t; int material_id; v3 normal, origin; } hit_t;   /* src/def.h:71-76 */
static inline hit_t no_hit(void) {                                       /* src/def.h:78-83 */
    hit_t h = {MAX_DIST + 1e1f, -1, {0, 0, 0}, {0, 0, 0}};
    return h;
}

typedef struct {
    const sbx_params* p;
    v2 res;
    float time;
    float mouse[4];
} uniforms_t;

/* ---- util.h ----------------------------------------------------------------------------------- */
static ray_t get_primary_ray(v3 cam_local_point, v3 cam_origin, v3 cam_look_at) {   /* src/util.h:5-20 */
    v3 fwd = norm3(sub3(cam_look_at, cam_origin));
    v3 up = V3(0, 1, 0);
    v3 right = cross3(up, fwd);
    up = cross3(fwd, right);
    ray_t r;
    r.origin = cam_origin;
    r.direction = norm3(add3(add3(fwd, scale3(up, cam_local_point.y)), scale3(right, cam_local_point.x)));
    return r;
}
static m3 rotate_around_z(float deg) {   /* src/util.h:45-52 */
    float a = g_radians(deg), s = m_sin(a), c = m_cos(a);
    return M3(c, -s, 0, s, c, 0, 0, 0, 1);
}
static m3 rotate_around_y(float deg) {   /* src/util.h:54-61 */
    float a = g_radians(deg), s = m_sin(a), c = m_cos(a);
    return M3(c, 0, s, 0, 1, 0, -s, 0, c);
}
static m3 rotate_around_x(float deg) {   /* src/util.h:63-69 */
    float a = g_radians(deg), s = m_sin(a), c = m_cos(a);
    return M3(1, 0, 0, 0, c, -s, 0, s, c);
}
static v3 linear_to_srgb(v3 color) {     /* src/util.h:72-77 */
    const float p = 1.0f / 2.2f;
    return V3(m_pow(color.x, p), m_pow(color.y, p), m_pow(color.z, p));
}
static v3 faceforward3(v3 N, v3 I, v3 Nref) { return dot3(Nref, I) < 0.0f ? N : neg3(N); }   /* src/util.h:86-92 */
static float band(float start, float peak, float end, float t) {                              /* src/util.h:103-112 */
    return g_smoothstep(start, peak, t) * (1.0f - g_smoothstep(peak, end, t));
}

/* ---- intersect.h ------------------------------------------------------------------------------ */
static void intersect_sphere(ray_t ray, sphere_t sphere, hit_t* hit) {   /* src/intersect.h:7-33 */
    v3 rc = sub3(sphere.origin, ray.origin);
    float radius2 = sphere.radius * sphere.radius;
    float tca = dot3(rc, ray.direction);
    if (tca < 0.0f) return;
    float d2 = dot3(rc, rc) - tca * tca;
    if (d2 > radius2) return;
    float thc = m_sqrt(radius2 - d2);
    float t0 = tca - thc;
    float t1 = tca + thc;
    if (t0 < 0.0f) t0 = t1;
    if (t0 > hit->t) return;
    v3 impact = add3(ray.origin, scale3(ray.direction, t0));
    hit->t = t0;
    hit->material_id = sphere.material;
    hit->origin = impact;
    hit->normal = divs3(sub3(impact, sphere.origin), sphere.radius);
}
static void intersect_plane(ray_t ray, plane_t p, hit_t* hit) {          /* src/intersect.h:61-77 */
    float denom = dot3(p.direction, ray.direction);
    if (denom < 1e-6f) return;
    v3 P0 = V3(p.distance, p.distance, p.distance);
    float t = dot3(sub3(P0, ray.origin), p.direction) / denom;
    if (t < 0.0f || t > hit->t) return;
    hit->t = t;
    hit->material_id = p.material;
    hit->origin = add3(ray.origin, scale3(ray.direction, t));
    hit->normal = faceforward3(p.direction, ray.direction, p.direction);
}

/* ---- noise_iq.h + fbm.h ----------------------------------------------------------------------- */
static inline float hash1(float n) { return g_fract(m_sin(n) * 753.5453123f); }   /* src/noise_iq.h:5-9 */
static float noise_iq(v3 x) {                                                      /* src/noise_iq.h:11-23 */
    v3 p = V3(floorf(x.x), floorf(x.y), floorf(x.z));
    v3 f = V3(g_fract(x.x), g_fract(x.y), g_fract(x.z));
    f = V3(f.x * f.x * (3.0f - 2.0f * f.x), f.y * f.y * (3.0f - 2.0f * f.y), f.z * f.z * (3.0f - 2.0f * f.z));
    float n = p.x + p.y * 157.0f + 113.0f * p.z;
    return g_mix(g_mix(g_mix(hash1(n + 0.0f), hash1(n + 1.0f), f.x),
                       g_mix(hash1(n + 157.0f), hash1(n + 158.0f), f.x), f.y),
                 g_mix(g_mix(hash1(n + 113.0f), hash1(n + 114.0f), f.x),
                       g_mix(hash1(n + 270.0f), hash1(n + 271.0f), f.x), f.y), f.z);
}
enum { BASIS_NOISE, BASIS_ABS, BASIS_RIDGE };
/* DECL_FBM_FUNC (src/fbm.h:6) for the three basis expressions the apps use */
static float fbm(v3 pos, int octaves, int basis, float lacunarity, float init_gain, float gain) {
    v3 p = pos;
    float H = init_gain, t = 0.0f;
    for (int i = 0; i < octaves; i++) {
        float b = noise_iq(p);
        if (basis == BASIS_ABS) b = fabsf(b * 2.0f - 1.0f);                 /* src/app_planet.h:65 */
        else if (basis == BASIS_RIDGE) b = 1.0f - fabsf(b * 2.0f - 1.0f);   /* src/app_planet.h:167 */
        t += b * H;
        p = scale3(p, lacunarity);
        H *= gain;
    }
    return t;
}

/* ---- volumetric.h ----------------------------------------------------------------------------- */
static float rayleigh_phase(float mu) { return 3.0f * (1.0f + mu * mu) / (16.0f * PI_F); }   /* :13-20 */
static float hg_phase(float mu, float g) {                                                   /* :27-33 */
    return (1.0f - g * g) / ((4.0f + PI_F) * m_pow(1.0f + g * g - 2.0f * g * mu, 1.5f));
}
typedef struct { v3 origin, pos; float height, transmittance; v3 radiance; float alpha; } volume_t;   /* :47-54 */
static volume_t construct_volume(v3 origin) {                                                /* :56-68 */
    volume_t v = {origin, origin, 0.0f, 1.0f, {0, 0, 0}, 0.0f};
    return v;
}

/* the shared mainImage body, src/main.h:6-53 */
typedef v3 (*render_fn)(const uniforms_t*, ray_t, v3);
static void main_image(const uniforms_t* u, render_fn render, v3 eye, v3 look_at, float fov, float fx, float fy,
                       float out[4]) {
    v2 aspect = V2(u->res.x / u->res.y, 1.0f);
    v2 ndc = V2(fx / u->res.x, fy / u->res.y);
    v3 point_cam = V3((2.0f * ndc.x - 1.0f) * aspect.x * fov, (2.0f * ndc.y - 1.0f) * aspect.y * fov, -1.0f);
    ray_t ray = get_primary_ray(point_cam, eye, look_at);
    v3 color = linear_to_srgb(render(u, ray, point_cam));
    out[0] = color.x; out[1] = color.y; out[2] = color.z; out[3] = 1.0f;
}

/* ================================= APP_CLOUDS (src/app_clouds.h) ============================== */
typedef struct { v3 wind_dir, sun_dir, sun_color; } clouds_vecs;
static clouds_vecs clouds_uniform_vecs(const sbx_params* p) {
    clouds_vecs v = {{p->wind_dir[0], p->wind_dir[1], p->wind_dir[2]},
                     {p->sun_dir[0], p->sun_dir[1], p->sun_dir[2]},
                     {p->sun_color[0], p->sun_color[1], p->sun_color[2]}};
    return v;
}
static v3 clouds_sky(const uniforms_t* u, v3 dir) {                 /* :36-46 */
    clouds_vecs cv = clouds_uniform_vecs(u->p);
    float sun_amount = g_max(dot3(dir, cv.sun_dir), 0.0f);
    v3 sky = mix3(V3(.0f, .1f, .4f), V3(.3f, .6f, .8f), 1.0f - dir.y);
    sky = add3(sky, scale3(cv.sun_color, g_min(m_pow(sun_amount, 1500.0f) * 5.0f, 1.0f)));
    sky = add3(sky, scale3(cv.sun_color, g_min(m_pow(sun_amount, 10.0f) * .6f, 1.0f)));
    return abs3(sky);
}
static float clouds_density(const uniforms_t* u, v3 pos_in) {       /* :62-86 */
    v3 pos = scale3(pos_in, .001f);
    float shape = fbm(scale3(pos, 2.03f), 4, BASIS_NOISE, 2.64f, .5f, .5f);
    const float cov = 1.0f - u->p->cld_coverage;
    return shape * g_smoothstep(cov, cov + .0135f, shape);
}
static float clouds_illuminate(const uniforms_t* u, v3 origin, v3 V, v3 L) {   /* :91-123 */
    const sbx_params* p = u->p;
    const float dt = p->cld_thick / (float)p->cld_march_steps;
    volume_t vol = construct_volume(origin);
    vol.pos = add3(vol.pos, scale3(L, dt));
    for (int i = 0; i < p->illum_march_steps; i++) {
        vol.height = (float)i / (float)p->illum_march_steps;
        float density = clouds_density(u, vol.pos);
        vol.transmittance *= m_exp(-density * p->sigma_scattering * dt);
        vol.pos = add3(vol.pos, scale3(L, dt));
    }
    float luminance = vol.transmittance;
    return luminance * p->sun_power * hg_phase(g_clamp(dot3(L, V), 0.0f, 1.0f), .2f);
}
static void clouds_integrate(const uniforms_t* u, volume_t* vol, v3 V, v3 L, float density, float dt) {   /* :125-148 */
    const sbx_params* p = u->p;
    if (density < .005f) return;
    float T_i = m_exp(-density * p->sigma_scattering * dt);
    vol->transmittance *= T_i;
    float add = (density * p->sigma_scattering) * clouds_illuminate(u, vol->pos, V, L) * vol->transmittance * dt;
    vol->radiance = adds3(vol->radiance, add);
    vol->alpha += (1.0f - T_i) * (1.0f - vol->alpha);
}
static v3 clouds_render(const uniforms_t* u, ray_t eye, v3 point_cam) {       /* :153-218 */
    (void)point_cam;
    const sbx_params* p = u->p;
    clouds_vecs cv = clouds_uniform_vecs(p);
    v3 sky = clouds_sky(u, eye.direction);
    if (dot3(eye.direction, V3(0, 1, 0)) < 0.05f) return sky;

    v3 projection = divs3(eye.direction, eye.direction.y);
    v3 origin = add3(eye.origin, scale3(projection, 150.0f));
    origin = add3(origin, scale3(scale3(cv.wind_dir, u->time), 1.0f / .001f));
    volume_t cloud = construct_volume(origin);
    float t = 0.0f;
    const float dt = p->cld_thick / (float)p->cld_march_steps;
    for (int i = 0; i < p->cld_march_steps; i++) {
        cloud.height = (float)i / (float)p->cld_march_steps;
        cloud.pos = add3(cloud.origin, rscale3(t, projection));
        t += dt;
        float density = clouds_density(u, cloud.pos);
        clouds_integrate(u, &cloud, eye.direction, cv.sun_dir, density, dt);
        if (cloud.alpha > .999f) break;
    }
    float cutoff = dot3(eye.direction, V3(0, 1, 0));
    float a = cloud.alpha * g_smoothstep(.0f, .2f, cutoff);
    return abs3(mix3(sky, cloud.radiance, a));
}
static void clouds_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {
    v3 eye = V3(0, -.5f, 0);                                         /* :23-30 */
    float angle = u->mouse[0] * .5f;
    v3 look_at = mat_vec(rotate_around_y(angle), V3(0, 0, -1));
    main_image(u, clouds_render, eye, look_at, 1.0f, fx, fy, out);   /* FOV 1. (:220) */
}

/* ========================= APP_CLOUDS with USE_NOISE_TEX (src/app_clouds.h:8-9) ================ *
 * The HLSL-only branch of the cloud app: the density comes from two 3-D noise textures (u_tex_noise, u_tex_noise_2:
 * :51-55) sampled with u_sampler0 -- D3D11_FILTER_MIN_MAG_MIP_LINEAR, WRAP addressing on u, v and w
 * (util/hlsltoy/src/hlsltoy.cpp:244-249) -- at LOD 0 (`SampleLevel(u_sampler0, pos, 0).r`, :69, :77).
 *
 * PARITY UNPINNED: the reference has no C++ statement of that sampler (the C++ build cannot compile this branch), no
 * texture files and no output images, and hardware filtering is only specified up to its fixed-point precision.  This
 * restatement therefore DEFINES the sampler, following the D3D11.3 functional specification's description of linear
 * filtering (7.18.8; D3D11_SUBTEXEL_FRACTIONAL_BIT_COUNT = 8):
 *     per axis:  u' = u - floor(u)                        (WRAP)
 *                t  = u' * N - 0.5                         (texel space, texel centres at i + 0.5)
 *                i0 = floor(t),  f = t - i0,  i1 = i0 + 1  (both wrapped into [0, N))
 *                w  = floor(f * 256 + 0.5) / 256           (8-bit sub-texel weight)
 *     value = lerp_z(lerp_y(lerp_x(...)))  with lerp(a, b, w) = a * (1 - w) + b * w in fp32, x innermost
 * Texels are the .r channel of size^3 R32G32B32A32_FLOAT volumes, x fastest (what ddsvolgen writes). */
static const float* g_noise_vol[2];
static int g_noise_size;
int sbxoracle_set_noise_volumes(const float* rgba_a, const float* rgba_b, int size) {
    if (!rgba_a || !rgba_b || size <= 0) return SBX_ERR_INVALID;
    g_noise_vol[0] = rgba_a; g_noise_vol[1] = rgba_b; g_noise_size = size;
    return SBX_OK;
}
typedef struct { int i0, i1; float w; } tex_axis;
static tex_axis tex_coord(float u, int n) {
    tex_axis a;
    float uw = u - floorf(u);
    float t = uw * (float)n - 0.5f;
    float fl = floorf(t);
    float f = t - fl;
    int i0 = (int)fl;
    a.w = floorf(f * 256.0f + 0.5f) / 256.0f;
    a.i0 = ((i0 % n) + n) % n;
    a.i1 = (a.i0 + 1) % n;
    return a;
}
static float tex_lerp(float a, float b, float w) { return a * (1.0f - w) + b * w; }
static float tex_sample_r(int which, v3 pos) {
    const float* vol = g_noise_vol[which];
    const int n = g_noise_size;
    tex_axis x = tex_coord(pos.x, n), y = tex_coord(pos.y, n), z = tex_coord(pos.z, n);
#define TEXEL(ix, iy, iz) vol[(((size_t)(iz) * n + (iy)) * n + (ix)) * 4]
    float c00 = tex_lerp(TEXEL(x.i0, y.i0, z.i0), TEXEL(x.i1, y.i0, z.i0), x.w);
    float c10 = tex_lerp(TEXEL(x.i0, y.i1, z.i0), TEXEL(x.i1, y.i1, z.i0), x.w);
    float c01 = tex_lerp(TEXEL(x.i0, y.i0, z.i1), TEXEL(x.i1, y.i0, z.i1), x.w);
    float c11 = tex_lerp(TEXEL(x.i0, y.i1, z.i1), TEXEL(x.i1, y.i1, z.i1), x.w);
#undef TEXEL
    return tex_lerp(tex_lerp(c00, c10, y.w), tex_lerp(c01, c11, y.w), z.w);
}
/* test hook: one sample of texture `which` at pos (the sampler rule above, as the frames use it) */
float sbxoracle_sample_noise(int which, float x, float y, float z) {
    if (which < 0 || which > 1 || !g_noise_vol[which]) return 0.0f;
    return tex_sample_r(which, V3(x, y, z));
}
static float g_remap(float v, float omin, float omax, float nmin, float nmax) {   /* src/util.h:127-138 */
    return nmin + (((v - omin) / (omax - omin)) * (nmax - nmin));
}
static float clouds_tex_density(const uniforms_t* u, v3 pos_in, float height) {   /* :62-86, USE_NOISE_TEX */
    v3 pos = scale3(pos_in, .001f);
    float shape = tex_sample_r(0, pos);
    float w = tex_sample_r(1, pos);
    float ww = w * (1.0f - height) + (1.0f - w) * height;            /* mix(w, 1. - w, height) */
    shape = g_remap(shape, ww * .7f, 1.0f, 0.0f, 1.0f);
    const float cov = 1.0f - u->p->cld_coverage;
    return shape * g_smoothstep(cov, cov + .0135f, shape);
}
static float clouds_tex_illuminate(const uniforms_t* u, v3 origin, v3 V, v3 L) {   /* :91-123 */
    const sbx_params* p = u->p;
    const float dt = p->cld_thick / (float)p->cld_march_steps;
    volume_t vol = construct_volume(origin);
    vol.pos = add3(vol.pos, scale3(L, dt));
    for (int i = 0; i < p->illum_march_steps; i++) {
        vol.height = (float)i / (float)p->illum_march_steps;
        float density = clouds_tex_density(u, vol.pos, vol.height);
        vol.transmittance *= m_exp(-density * p->sigma_scattering * dt);
        vol.pos = add3(vol.pos, scale3(L, dt));
    }
    return vol.transmittance * p->sun_power * hg_phase(g_clamp(dot3(L, V), 0.0f, 1.0f), .2f);
}
static v3 clouds_tex_render(const uniforms_t* u, ray_t eye, v3 point_cam) {       /* :153-218 */
    (void)point_cam;
    const sbx_params* p = u->p;
    clouds_vecs cv = clouds_uniform_vecs(p);
    v3 sky = clouds_sky(u, eye.direction);
    if (dot3(eye.direction, V3(0, 1, 0)) < 0.05f) return sky;
    v3 projection = divs3(eye.direction, eye.direction.y);
    v3 origin = add3(eye.origin, scale3(projection, 150.0f));
    origin = add3(origin, scale3(scale3(cv.wind_dir, u->time), 1.0f / .001f));
    volume_t cloud = construct_volume(origin);
    float t = 0.0f;
    const float dt = p->cld_thick / (float)p->cld_march_steps;
    for (int i = 0; i < p->cld_march_steps; i++) {
        cloud.height = (float)i / (float)p->cld_march_steps;
        cloud.pos = add3(cloud.origin, rscale3(t, projection));
        t += dt;
        float density = clouds_tex_density(u, cloud.pos, cloud.height);
        if (!(density < .005f)) {                                    /* integrate_volume, :125-148 */
            float T_i = m_exp(-density * p->sigma_scattering * dt);
            cloud.transmittance *= T_i;
            float add = (density * p->sigma_scattering) * clouds_tex_illuminate(u, cloud.pos, eye.direction, cv.sun_dir) *
                        cloud.transmittance * dt;
            cloud.radiance = adds3(cloud.radiance, add);
            cloud.alpha += (1.0f - T_i) * (1.0f - cloud.alpha);
        }
        if (cloud.alpha > .999f) break;
    }
    float cutoff = dot3(eye.direction, V3(0, 1, 0));
    float a = cloud.alpha * g_smoothstep(.0f, .2f, cutoff);
    return abs3(mix3(sky, cloud.radiance, a));
}
static void clouds_tex_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {
    v3 eye = V3(0, -.5f, 0);
    float angle = u->mouse[0] * .5f;
    v3 look_at = mat_vec(rotate_around_y(angle), V3(0, 0, -1));
    if (!g_noise_vol[0]) { out[0] = out[1] = out[2] = out[3] = 0.0f; return; }
    main_image(u, clouds_tex_render, eye, look_at, 1.0f, fx, fy, out);
}

/* ============================== APP_ATMOSPHERE (src/app_atmosphere.h) ========================= */
#define ATM_EARTH_RADIUS 6360e3f
#define ATM_RADIUS 6420e3f
static int atm_isect_sphere(ray_t ray, sphere_t sphere, float* t0, float* t1) {   /* :15-26 */
    v3 rc = sub3(sphere.origin, ray.origin);
    float radius2 = sphere.radius * sphere.radius;
    float tca = dot3(rc, ray.direction);
    float d2 = dot3(rc, rc) - tca * tca;
    float thc = m_sqrt(radius2 - d2);
    *t0 = tca - thc;
    *t1 = tca + thc;
    return d2 < radius2;
}
static int atm_sun_light(ray_t ray, float* odR, float* odM) {                      /* :50-76 */
    const float hR = 7994.0f, hM = 1200.0f;
    const sphere_t atmosphere = {{0, 0, 0}, ATM_RADIUS, 0};
    float t0, t1;
    atm_isect_sphere(ray, atmosphere, &t0, &t1);
    float march_pos = 0.0f;
    float march_step = t1 / (float)8;
    for (int i = 0; i < 8; i++) {
        v3 s = add3(ray.origin, scale3(ray.direction, march_pos + 0.5f * march_step));
        float height = len3(s) - ATM_EARTH_RADIUS;
        if (height < 0.0f) return 0;
        *odR += m_exp(-height / hR) * march_step;
        *odM += m_exp(-height / hM) * march_step;
        march_pos += march_step;
    }
    return 1;
}
static v3 atm_incident_light(ray_t ray, v3 sun_dir) {                               /* :78-160 */
    const v3 betaR = {5.5e-6f, 13.0e-6f, 22.4e-6f}, betaM = {21e-6f, 21e-6f, 21e-6f};
    const float hR = 7994.0f, hM = 1200.0f, sun_power = 20.0f;
    const sphere_t atmosphere = {{0, 0, 0}, ATM_RADIUS, 0};
    float t0, t1;
    if (!atm_isect_sphere(ray, atmosphere, &t0, &t1)) return V3(0, 0, 0);
    float march_step = t1 / (float)16;
    float mu = dot3(ray.direction, sun_dir);
    float phaseR = rayleigh_phase(mu);
    float phaseM = hg_phase(mu, .76f);
    float odR = 0.0f, odM = 0.0f;
    v3 sumR = {0, 0, 0}, sumM = {0, 0, 0};
    float march_pos = 0.0f;
    for (int i = 0; i < 16; i++) {
        v3 s = add3(ray.origin, scale3(ray.direction, march_pos + 0.5f * march_step));
        float height = len3(s) - ATM_EARTH_RADIUS;
        float hr = m_exp(-height / hR) * march_step;
        float hm = m_exp(-height / hM) * march_step;
        odR += hr;
        odM += hm;
        ray_t light_ray = {s, sun_dir};
        float olR = 0.0f, olM = 0.0f;
        if (atm_sun_light(light_ray, &olR, &olM)) {
            v3 tau = add3(scale3(betaR, odR + olR), scale3(scale3(betaM, 1.1f), odM + olM));
            v3 att = V3(m_exp(-tau.x), m_exp(-tau.y), m_exp(-tau.z));
            sumR = add3(sumR, rscale3(hr, att));
            sumM = add3(sumM, rscale3(hm, att));
        }
        march_pos += march_step;
    }
    return rscale3(sun_power, add3(mul3(scale3(sumR, phaseR), betaR), mul3(scale3(sumM, phaseM), betaM)));
}
static v3 atm_render(const uniforms_t* u, ray_t eye, v3 point_cam) {                /* :164-228, FROM_SPACE */
    (void)eye;
    /* setup_scene: the per-pixel sun_dir = (0,1,0) * rot  (:177-181) */
    m3 rot = rotate_around_x(-fabsf(m_sin(u->time / 2.0f)) * 90.0f);
    v3 sun_dir = vec_mat(V3(0, 1, 0), rot);
    v3 p = point_cam;
    float z2 = p.x * p.x + p.y * p.y;
    float phi = m_atan2(p.y, p.x);
    float theta = m_acos(1.0f - z2);
    float st = m_sin(theta), cp = m_cos(phi), ct = m_cos(theta), st2 = m_sin(theta), sp = m_sin(phi);
    v3 dir = V3(st * cp, ct, st2 * sp);
    ray_t ray = {{0, ATM_EARTH_RADIUS + 1.0f, 0}, dir};
    return atm_incident_light(ray, sun_dir);
}
static void atmosphere_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {
    main_image(u, atm_render, V3(0, 0, 0), V3(0, 1, 0), 1.0f, fx, fy, out);         /* :164-175, FOV 1. */
}

/* ================================= APP_PLANET (src/app_planet.h) ============================== */
#define PL_MAX_HEIGHT .4f
#define PL_MAX_RAY_DIST (PL_MAX_HEIGHT * 4.0f)
#define PL_TERR_EPS .005f
static const sphere_t pl_planet = {{0, 0, 0}, 1.0f, 0};                             /* :17-19 */
static v3 planet_background(ray_t eye) {                                            /* :23-41 */
    const v3 sun_color = {1.0f, .9f, .55f};
    float sun_amount = g_clamp(dot3(eye.direction, V3(0, 0, 1)), 0.0f, 1.0f);
    v3 sky = mix3(V3(.0f, .05f, .2f), V3(.15f, .3f, .4f), 1.0f - eye.direction.y);
    sky = add3(sky, scale3(sun_color, g_clamp(m_pow(sun_amount, 30.0f) * 5.0f, 0.0f, 1.0f)));
    sky = add3(sky, scale3(sun_color, g_clamp(m_pow(sun_amount, 10.0f) * .6f, 0.0f, 1.0f)));
    return abs3(sky);
}
static void planet_integrate(volume_t* vol, float density, float dt) {              /* :71-100 */
    float T_i = m_exp(-30.034f * density * dt);
    vol->transmittance *= T_i;
    float add = density * (m_exp(vol->height) / .055f) * vol->transmittance * dt;
    vol->radiance = adds3(vol->radiance, add);
    vol->alpha += (1.0f - T_i) * (1.0f - vol->alpha);
}
static void planet_clouds_map(volume_t* cloud, float t_step) {                      /* :102-119 */
    float dens = fbm(add3(scale3(cloud->pos, 3.2343f), V3(.35f, 13.35f, 2.67f)), 4, BASIS_ABS, 2.0276f, .5f, .5f);
    dens *= g_smoothstep(.29475675f, .29475675f + .0335f, dens);
    dens *= band(.2f, .35f, .65f, cloud->height);
    planet_integrate(cloud, dens, t_step);
}
static void planet_clouds_march(ray_t eye, volume_t* cloud, float max_travel, m3 rot) {   /* :121-141 */
    const int steps = 75;
    const float t_step = PL_MAX_RAY_DIST / (float)steps;
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        if (t > max_travel || cloud->alpha >= 1.0f) return;
        v3 o = add3(cloud->origin, rscale3(t, eye.direction));
        cloud->pos = mat_vec(rot, sub3(o, pl_planet.origin));
        cloud->height = (len3(cloud->pos) - pl_planet.radius) / PL_MAX_HEIGHT;
        t += t_step;
        planet_clouds_map(cloud, t_step);
    }
}
static void planet_shadow_march(v3 dir, volume_t* cloud, m3 rot) {                  /* :143-160 */
    const int steps = 5;
    const float t_step = PL_MAX_HEIGHT / (float)steps;
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        v3 o = add3(cloud->origin, rscale3(t, dir));
        cloud->pos = mat_vec(rot, sub3(o, pl_planet.origin));
        cloud->height = (len3(cloud->pos) - pl_planet.radius) / PL_MAX_HEIGHT;
        t += t_step;
        planet_clouds_map(cloud, t_step);
    }
}
static v2 planet_terrain_map(v3 pos, int octaves) {                                 /* :175-199 (3 or 7 octaves) */
    float h0 = fbm(scale3(pos, 2.0987f), octaves, BASIS_NOISE, 2.0244f, .454f, .454f);
    float n0 = g_smoothstep(.35f, 1.0f, h0);
    float h1 = fbm(add3(scale3(pos, 1.50987f), V3(1.9489f, 2.435f, .5483f)), octaves, BASIS_RIDGE, 2.0244f, .454f, .454f);
    float n1 = g_smoothstep(.6f, 1.0f, h1);
    float n = n0 + n1;
    return V2(len3(pos) - pl_planet.radius - n * PL_MAX_HEIGHT, n / PL_MAX_HEIGHT);
}
static v3 planet_terrain_normal(v3 p) {                                             /* :201-212 */
    const float e = 0.001f;
    v3 px = V3(e, 0, 0), py = V3(0, e, 0), pz = V3(0, 0, e);
    return norm3(V3(planet_terrain_map(add3(p, px), 7).x - planet_terrain_map(sub3(p, px), 7).x,
                    planet_terrain_map(add3(p, py), 7).x - planet_terrain_map(sub3(p, py), 7).x,
                    planet_terrain_map(add3(p, pz), 7).x - planet_terrain_map(sub3(p, pz), 7).x));
}
static v3 planet_lights(v3 L, v3 normal) {                                          /* :217-236 */
    v3 diffuse = {0, 0, 0};
    v3 c_L = {7, 5, 3};
    diffuse = add3(diffuse, rscale3(g_max(0.0f, dot3(L, normal)), c_L));
    float hemi = g_clamp(.25f + .5f * normal.y, .0f, 1.0f);
    diffuse = add3(diffuse, scale3(rscale3(hemi, V3(.4f, .6f, .8f)), .2f));
    float amb = g_clamp(.12f + .8f * g_max(0.0f, dot3(neg3(L), normal)), 0.0f, 1.0f);
    diffuse = add3(diffuse, rscale3(amb, V3(.4f, .5f, .6f)));
    return diffuse;
}
static v3 planet_illuminate(v3 pos, m3 local_xform, v2 df) {                        /* :238-298 */
    const v3 c_water = {.015f, .110f, .455f}, c_grass = {.086f, .132f, .018f}, c_beach = {.153f, .172f, .121f},
             c_rock = {.080f, .050f, .030f}, c_snow = {.600f, .600f, .600f};
    const float l_water = .05f, l_shore = .17f, l_grass = .211f, l_rock = .351f;
    float h = df.y;
    v3 w_normal = norm3(pos);
    v3 normal = planet_terrain_normal(pos);
    float N = dot3(normal, w_normal);
    float s = g_smoothstep(.4f, 1.0f, h);
    v3 rock = mix3(c_rock, c_snow, g_smoothstep(1.0f - .3f * s, 1.0f - .2f * s, N));
    v3 grass = mix3(c_grass, rock, g_smoothstep(l_grass, l_rock, h));
    v3 shoreline = mix3(c_beach, grass, g_smoothstep(l_shore, l_grass, h));
    v3 water = mix3(divs3(c_water, 2.0f), c_water, g_smoothstep(0.0f, l_water, h));
    v3 L = mat_vec(local_xform, norm3(V3(1, 1, 0)));
    shoreline = mul3(shoreline, planet_lights(L, normal));
    v3 ocean = mul3(planet_lights(L, w_normal), water);
    return mix3(ocean, shoreline, g_smoothstep(l_water, l_shore, h));
}
static v3 planet_render(const uniforms_t* u, ray_t eye, v3 point_cam) {             /* :303-367 */
    (void)point_cam;
    m3 rot_y = rotate_around_y(27.0f);
    m3 rot = mat_mat(rotate_around_x(u->time * -12.0f), rot_y);
    m3 rot_cloud = mat_mat(rotate_around_x(u->time * 8.0f), rot_y);
    sphere_t atmosphere = pl_planet;
    atmosphere.radius += PL_MAX_HEIGHT;
    hit_t hit = no_hit();
    intersect_sphere(eye, atmosphere, &hit);
    if (hit.material_id < 0) return planet_background(eye);

    float t = 0.0f;
    v2 df = V2(1, PL_MAX_HEIGHT);
    v3 pos = {0, 0, 0};
    float max_cld_ray_dist = PL_MAX_RAY_DIST;
    for (int i = 0; i < 120; i++) {
        if (t > PL_MAX_RAY_DIST) break;
        v3 o = add3(hit.origin, rscale3(t, eye.direction));
        pos = mat_vec(rot, sub3(o, pl_planet.origin));
        df = planet_terrain_map(pos, 3);
        if (df.x < PL_TERR_EPS) {
            max_cld_ray_dist = t;
            break;
        }
        t += df.x * .4567f;
    }
    volume_t cloud = construct_volume(hit.origin);
    planet_clouds_march(eye, &cloud, max_cld_ray_dist, rot_cloud);

    if (df.x < PL_TERR_EPS) {
        v3 c_terr = planet_illuminate(pos, rot, df);
        v3 c_cld = cloud.radiance;
        float alpha = cloud.alpha;
        pos = mat_vec(transpose3(rot), pos);
        cloud = construct_volume(pos);
        v3 local_up = norm3(pos);
        planet_shadow_march(local_up, &cloud, rot_cloud);
        float shadow = g_mix(.7f, 1.0f, g_step(cloud.alpha, 0.33f));
        return abs3(mix3(scale3(c_terr, shadow), c_cld, alpha));
    }
    return abs3(mix3(planet_background(eye), cloud.radiance, cloud.alpha));
}
static void planet_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {
    const float fov = m_tan(g_radians(30.0f));                                      /* :369 */
    main_image(u, planet_render, V3(0, 0, -2.5f), V3(0, 0, 2), fov, fx, fy, out);   /* :45-56 */
}

/* =============================== APP_RAYTRACER (src/app_raytracer.h) ========================== */
typedef struct { v3 base_color; float metallic, roughness, ior, reflectivity, translucency; } material_t;   /* src/material.h:5-12 */
typedef struct { int type; v3 L, color; } light_t;                                                          /* src/light.h:8-12 */
typedef struct {
    material_t materials[8];
    light_t light0;
    v3 ambient;
    plane_t planes[6];
    sphere_t spheres[3];
} rt_scene;
#define CB_DIST 2.0f
static void rt_material(material_t* m, v3 diffuse, float metallic, float roughness) {   /* src/cornell_box.h:14-26 */
    m->base_color = diffuse; m->metallic = metallic; m->roughness = roughness;
    m->ior = 1.0f; m->reflectivity = 0.0f; m->translucency = 0.0f;
}
static void rt_setup_scene(const uniforms_t* u, rt_scene* s) {   /* src/app_raytracer.h:18-36 + src/cornell_box.h:39-87 */
    memset(s, 0, sizeof *s);
    s->ambient = V3(.01f, .01f, .01f);                           /* src/light.h:16 */
    rt_material(&s->materials[0], V3(1, 1, 1), 0.0f, 0.0f);      /* mat_debug */
    rt_material(&s->materials[1], V3(0.7913f, 0.7913f, 0.7913f), .0f, .5f);
    rt_material(&s->materials[2], V3(0.6795f, 0.0612f, 0.0529f), 0.0f, .5f);
    rt_material(&s->materials[3], V3(0.1878f, 0.1274f, 0.4287f), 0.0f, .5f);
    rt_material(&s->materials[4], V3(0.95f, 0.64f, 0.54f), 1.0f, .1f);
    s->materials[4].reflectivity = 1.0f;
    rt_material(&s->materials[5], V3(1.0f, 0.77f, 0.345f), 1.0f, .05f);
    s->materials[5].reflectivity = 1.0f;
    s->materials[5].ior = 1.333f;
    /* planes in ARRAY order: ground, behind, front, ceiling, left, right (cornell_box.h:62-74) */
    plane_t ground = {{0, -1, 0}, 0.0f, 1}, behind = {{0, 0, -1}, -CB_DIST, 1}, front = {{0, 0, 1}, CB_DIST, 1},
            ceiling = {{0, 1, 0}, 2.0f * CB_DIST, 1}, left = {{1, 0, 0}, CB_DIST, 2}, right = {{-1, 0, 0}, -CB_DIST, 3};
    s->planes[0] = ground; s->planes[1] = behind; s->planes[2] = front;
    s->planes[3] = ceiling; s->planes[4] = left; s->planes[5] = right;
    sphere_t lamp = {{0, 2.5f * CB_DIST + 0.4f, 0}, 1.5f, 0}, ball_l = {{0.75f, 1, -0.75f}, 0.75f, 4},
             ball_r = {{-0.75f, 0.75f, 0.75f}, 0.75f, 5};
    s->spheres[0] = lamp; s->spheres[1] = ball_l; s->spheres[2] = ball_r;
    s->light0.type = 1;
    s->light0.L = V3(0, 2.0f * CB_DIST - 0.2f, 0);
    s->light0.color = V3(1, 1, 1);
    /* animation, app_raytracer.h:29-35 */
    float sn = m_sin(u->time), cs = m_cos(u->time);
    s->spheres[1].origin = add3(s->spheres[1].origin, V3(0, fabsf(sn), cs + 1.0f));
    s->spheres[2].origin.z = 0.0f;
    s->light0.L.z = 1.5f;
}
static float fresnel_factor(float n1, float n2, float VdotH) {                    /* src/util_optics.h:5-14 */
    float Rn = (n1 - n2) / (n1 + n2);
    float R0 = Rn * Rn;
    float F = 1.0f - VdotH;
    return R0 + (1.0f - R0) * (F * F * F * F * F);
}
static v3 reflect3(v3 incident, v3 normal) {                                      /* src/util_optics.h:17-22 */
    return sub3(incident, rscale3(2.0f * dot3(normal, incident), normal));
}
static v3 cook_torrance(v3 V, v3 L, const hit_t* hit, const material_t* mat) {   /* src/light.h:64-92 */
    v3 H = norm3(add3(L, V));
    float NdotL = dot3(hit->normal, L);
    float NdotH = dot3(hit->normal, H);
    float NdotV = dot3(hit->normal, V);
    float VdotH = dot3(V, H);
    float geo_a = (2.0f * NdotH * NdotV) / VdotH;
    float geo_b = (2.0f * NdotH * NdotL) / VdotH;
    float geo_term = g_min(1.0f, g_min(geo_a, geo_b));
    float rough_sq = mat->roughness * mat->roughness;
    float rough_a = 1.0f / (rough_sq * NdotH * NdotH * NdotH * NdotH);
    float rough_exp = (NdotH * NdotH - 1.0f) / (rough_sq * NdotH * NdotH);
    float rough_term = rough_a * m_exp(rough_exp);
    float fresnel_term = fresnel_factor(1.0f, mat->ior, VdotH);
    float specular = (geo_term * rough_term * fresnel_term) / (PI_F * NdotV * NdotL);
    return rscale3(g_max(0.0f, NdotL), V3(specular + mat->base_color.x, specular + mat->base_color.y, specular + mat->base_color.z));
}
static v3 rt_illuminate(const rt_scene* s, v3 eye, const hit_t* hit) {            /* src/app_raytracer.h:46-68 */
    if (hit->material_id == 0) return s->materials[0].base_color;
    const material_t* mat = &s->materials[hit->material_id];
    v3 accum = s->ambient;
    v3 V = norm3(sub3(eye, hit->origin));
    v3 L = norm3(sub3(s->light0.L, hit->origin));                                 /* point light, src/light.h:18-27 */
    return add3(accum, cook_torrance(V, L, hit, mat));
}
static hit_t rt_trace(const rt_scene* s, ray_t ray, int mat_to_ignore) {          /* src/app_raytracer.h:70-86 */
    hit_t hit = no_hit();
    for (int i = 0; i < 6; ++i) intersect_plane(ray, s->planes[i], &hit);
    for (int i = 0; i < 3; ++i)
        if (s->spheres[i].material != mat_to_ignore) intersect_sphere(ray, s->spheres[i], &hit);
    return hit;
}
static v3 rt_render(const uniforms_t* u, ray_t primary_ray, v3 point_cam) {       /* src/app_raytracer.h:88-136 */
    (void)point_cam;
    rt_scene s;
    rt_setup_scene(u, &s);
    v3 color = {0, 0, 0}, accum = {1, 1, 1};
    ray_t ray = primary_ray;
    for (int i = 0; i < 2; i++) {
        hit_t hit = rt_trace(&s, ray, -1);
        if (hit.t >= MAX_DIST) {
            color = add3(color, mul3(accum, V3(0, 0, 0)));
            break;
        }
        float f = fresnel_factor(1.0f, 1.0f, dot3(hit.normal, neg3(ray.direction)));
        color = add3(color, mul3(rscale3(1.0f - f, accum), rt_illuminate(&s, primary_ray.origin, &hit)));
        if (i == 0) {   /* shadow ray */
            v3 shadow_line = sub3(s.light0.L, hit.origin);
            v3 shadow_dir = norm3(shadow_line);
            ray_t shadow_trace = {add3(hit.origin, scale3(shadow_dir, BIAS)), shadow_dir};
            hit_t shadow_hit = rt_trace(&s, shadow_trace, 0);
            if (shadow_hit.t < len3(shadow_line)) color = scale3(color, 0.1f);
        }
        const material_t* mat = &s.materials[hit.material_id];
        if (mat->reflectivity > 0.0f) {
            accum = scale3(accum, f);
            v3 reflect_dir = norm3(reflect3(hit.normal, ray.direction));   /* (sic) argument order of the reference, :127 */
            ray.origin = add3(hit.origin, scale3(reflect_dir, BIAS));
            ray.direction = reflect_dir;
        } else {
            break;
        }
    }
    return color;
}
static void raytracer_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {   /* :38-44, :137 */
    v2 mouse = V2(0, 0);
    if (!(u->mouse[0] < BIAS)) {
        v2 q = V2(u->res.x / u->mouse[0], u->res.y / u->mouse[1]);
        mouse = V2(2.0f * q.x - 1.0f, 2.0f * q.y - 1.0f);
    }
    m3 rot_y = rotate_around_y(mouse.x * 30.0f);
    v3 eye = mat_vec(rot_y, V3(0, CB_DIST, 2.333f * CB_DIST));
    v3 look_at = V3(0, CB_DIST, 0);
    const float fov = m_tan(g_radians(30.0f));
    main_image(u, rt_render, eye, look_at, fov, fx, fy, out);
}

/* ==================================== APP_EGG (src/app_egg.h) ================================= */
static float op_blend(float a, float b, float k) {                               /* src/sdf.h:38-47 */
    float h = g_clamp(0.5f + 0.5f * (b - a) / k, 0.0f, 1.0f);
    return g_mix(b, a, h) - k * h * (1.0f - h);
}
static inline v2 op_add2(v2 d1, v2 d2) { return d1.x < d2.x ? d1 : d2; }         /* src/sdf.h:5-11 */
static inline float sd_plane(v3 p, v3 n, float d) { return dot3(n, p) + d; }     /* src/sdf.h:49-57 */
static inline float sd_sphere(v3 p, float r) { return len3(p) - r; }             /* src/sdf.h:59-65 */
static float sd_torus(v3 p, float R, float r) {                                  /* src/sdf.h:75-83 */
    return len2(V2(len2(V2(p.x, p.y)) - R, p.z)) - r;
}
static float sd_cylinder(v3 P, v3 P0, v3 P1, float R) {                          /* src/sdf.h:95-109 */
    v3 dir = norm3(sub3(P1, P0));
    float dist = len3(cross3(dir, sub3(P, P0)));
    float plane_1 = sd_plane(P, dir, len3(P1));
    float plane_2 = sd_plane(P, neg3(dir), -len3(P0));
    return g_max(g_max(dist, -plane_1), -plane_2) - R;                           /* op_sub twice, src/sdf.h:20-28 */
}
static inline float det2(v2 a, v2 b) { return a.x * b.y - b.x * a.y; }           /* src/sdf.h:114-119 */
static v3 bezier_closest(v2 b0, v2 b1, v2 b2) {                                  /* src/sdf.h:120-139 */
    float a = det2(b0, b2);
    float b = 2.0f * det2(b1, b0);
    float d = 2.0f * det2(b2, b1);
    float f = b * d - a * a;
    v2 d21 = sub2(b2, b1), d10 = sub2(b1, b0), d20 = sub2(b2, b0);
    v2 gf = rscale2(2.0f, add2(add2(rscale2(b, d21), rscale2(d, d10)), rscale2(a, d20)));
    gf = V2(gf.y, -gf.x);
    v2 pp = divs2(rscale2(-f, gf), dot2(gf, gf));
    v2 d0p = sub2(b0, pp);
    float ap = det2(d0p, d20);
    float bp = 2.0f * det2(d10, d0p);
    float t = g_clamp((ap + bp) / (2.0f * a + b + d), 0.0f, 1.0f);
    v2 q = mix2(mix2(b0, b1, t), mix2(b1, b2, t), t);
    return V3(q.x, q.y, t);
}
static v2 sd_bezier(v3 a, v3 b, v3 c, v3 p, float thickness) {                   /* src/sdf.h:140-159 */
    v3 w = norm3(cross3(sub3(c, b), sub3(a, b)));
    v3 u = norm3(sub3(c, b));
    v3 v = norm3(cross3(w, u));
    v2 a2 = V2(dot3(sub3(a, b), u), dot3(sub3(a, b), v));
    v2 b2 = V2(0, 0);
    v2 c2 = V2(dot3(sub3(c, b), u), dot3(sub3(c, b), v));
    v3 p3 = V3(dot3(sub3(p, b), u), dot3(sub3(p, b), v), dot3(sub3(p, b), w));
    v2 pxy = V2(p3.x, p3.y);
    v3 cp = bezier_closest(sub2(a2, pxy), sub2(b2, pxy), sub2(c2, pxy));
    return V2(0.85f * (m_sqrt(dot2(V2(cp.x, cp.y), V2(cp.x, cp.y)) + p3.z * p3.z) - thickness), cp.z);
}
static v3 ik_solver(v3 start, v3 goal_abs, float L1, float L2) {                 /* src/IK.h:5-52 */
    v3 goal = sub3(goal_abs, start);
    float G = len3(goal);
    float cos_theta = (L1 * L1 + G * G - L2 * L2) / (2.0f * L1 * G);
    float sin_theta = m_sqrt(1.0f - cos_theta * cos_theta);
    m3 rot = M3(cos_theta, -sin_theta, 0, sin_theta, cos_theta, 0, 0, 0, 1.0f);
    return add3(start, mat_vec(rot, scale3(norm3(goal), L1)));
}
static v2 egg_sdf(const uniforms_t* u, v3 P) {                                   /* src/app_egg.h:38-144 */
    v3 p = sub3(mat_vec(rotate_around_y(u->time * -100.0f), P), V3(0, 0.5f, 3.5f));
    const float material = 1.0f;   /* mat_egg */
    float egg_y = 0.65f;
    float egg_m = sd_sphere(sub3(p, V3(0, egg_y, 0)), 0.475f);
    float egg_b = sd_sphere(sub3(p, V3(0, egg_y - 0.45f, 0)), 0.25f);
    float egg_t = sd_sphere(sub3(p, V3(0, egg_y + 0.45f, 0)), 0.25f);
    float egg_1 = op_blend(egg_m, egg_b, .5f);
    float egg_2 = op_blend(egg_1, egg_t, .5f);
    v2 egg = V2(egg_2, material);

    v3 wheel_pos = V3(0, 1.2f, 0);
    float pedal_radius = 0.3f, pedal_speed = 400.0f, pedal_off = 0.2f;
    m3 rot_z = rotate_around_z(-u->time * pedal_speed);
    v3 left_foot_pos = add3(wheel_pos, mat_vec(rot_z, V3(0.0f, pedal_radius, pedal_off)));
    rot_z = rotate_around_z(-u->time * pedal_speed);
    v3 right_foot_pos = add3(wheel_pos, mat_vec(rot_z, V3(0.0f, -pedal_radius, -pedal_off)));

    v3 side = V3(0, 0, pedal_off);
    float femur = 0.8f, tibia = 0.75f, thick = .05f;
    v3 pelvis = add3(V3(0, 0.0f, 0), side);
    v3 knee_l = ik_solver(pelvis, left_foot_pos, femur, tibia);
    pelvis = sub3(V3(0, 0.0f, 0), side);
    v3 knee_r = ik_solver(pelvis, right_foot_pos, femur, tibia);

    v2 legs = op_add2(
        V2(sd_bezier(neg3(add3(V3(0, 0, 0), side)), neg3(knee_l), neg3(left_foot_pos), p, thick).x, material),
        V2(sd_bezier(neg3(sub3(V3(0, 0, 0), side)), neg3(knee_r), neg3(right_foot_pos), p, thick).x, material));

    v3 left_toe = norm3(V3(left_foot_pos.y - knee_l.y, knee_l.x - left_foot_pos.x, 0));
    v2 left_foot = V2(sd_cylinder(add3(p, left_foot_pos), V3(0, 0, 0), divs3(left_toe, 8.0f), thick), material);
    v3 right_toe = norm3(V3(right_foot_pos.y - knee_r.y, knee_r.x - right_foot_pos.x, 0));
    v2 right_foot = V2(sd_cylinder(add3(p, right_foot_pos), V3(0, 0, 0), divs3(right_toe, 8.0f), thick), material);
    v2 feet = op_add2(left_foot, right_foot);

    v2 bike = V2(sd_torus(add3(p, wheel_pos), 1.0f, .03f), 2.0f);                   /* mat_bike */
    v2 ground = V2(sd_plane(P, V3(0.0f, 1.0f, 0.0f), wheel_pos.y + 0.5f), 3.0f);    /* mat_ground */
    v2 s1 = op_add2(feet, bike);
    v2 s2 = op_add2(egg, s1);
    v2 s3 = op_add2(legs, s2);
    return op_add2(ground, s3);
}
static float egg_shadowmarch(const uniforms_t* u, ray_t ray) {                    /* :161-186 */
    const int steps = 20;
    const float end = 10.0f, penumbra_factor = 15.0f, darkest = 0.1f;
    float t = 0.0f, umbra = 1.0f;
    for (int i = 0; i < steps; i++) {
        v3 p = add3(ray.origin, scale3(ray.direction, t));
        v2 d = egg_sdf(u, p);
        if (t > end) break;
        if (d.x < 0.001f) return darkest;
        t += d.x;
        umbra = g_min(umbra, penumbra_factor * d.x / t);
    }
    return umbra;
}
typedef struct { float depth; } egg_state;   /* the per-pixel `_mutable(float) depth` (:188) */
static v3 egg_render_scene(const uniforms_t* u, egg_state* st, ray_t ray) {      /* :190-231 */
    const int steps = 80;
    const float end = 15.0f;
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        v3 p = add3(ray.origin, scale3(ray.direction, t));
        v2 d = egg_sdf(u, p);
        if (t > end) break;
        if (d.x < 0.001f) {
            int material_id = (int)d.y;
            if (material_id == 1 || material_id == 2) st->depth = g_max(st->depth, p.z);
            float s = 1.0f;
            if ((int)d.y == 3) {
                v3 sh_dir = V3(0, 1, 1);
                ray_t sh_ray = {add3(p, scale3(sh_dir, 0.05f)), sh_dir};
                s = egg_shadowmarch(u, sh_ray);
            }
            v3 c;                                                                 /* illuminate, :26-32 */
            if (material_id == 3) c = V3(13.0f / 255.0f, 104.0f / 255.0f, 0.0f / 255.0f);
            else if (material_id == 1) c = V3(0.9f, 0.95f, 0.95f);
            else if (material_id == 2) c = V3(.2f, .2f, .2f);
            else c = V3(1, 1, 1);
            return scale3(c, s);
        }
        t += d.x;
    }
    return V3(.1f, .1f, .7f);                                                     /* background, :11-14 */
}
static v3 egg_render(const uniforms_t* u, ray_t eye, v3 point_cam) {             /* :233-251 */
    egg_state st = {-MAX_DIST};
    v3 final_color = egg_render_scene(u, &st, eye);
    float bar_factor = 1.0f - g_smoothstep(0.0f, 0.01f, fabsf((fabsf(point_cam.x) - 0.6f)) - 0.05f);
    float depth_factor = 1.0f - g_step(1.0f, st.depth);
    final_color = mix3(final_color, V3(.6f, .6f, .6f), bar_factor * depth_factor);
    return abs3(final_color);
}
static void egg_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {   /* :20-24, FOV 1. */
    main_image(u, egg_render, V3(.0f, .25f, 5.25f), V3(.0f, .25f, .0f), 1.0f, fx, fy, out);
}

/* ================================== APP_SDF_AO (src/app_sdf_ao.h) ============================= */
static inline float sd_box(v3 p, v3 b) {                                          /* src/sdf.h:67-73 */
    return g_max(fabsf(p.x) - b.x, g_max(fabsf(p.y) - b.y, fabsf(p.z) - b.z));
}
static inline float sd_y_cylinder(v3 p, float r, float h) {                       /* src/sdf.h:85-93 */
    return g_max(len2(V2(p.x, p.z)) - r, fabsf(p.y) - h / 2.0f);
}
static inline float op_sub1(float d1, float d2) { return g_max(d1, -d2); }        /* src/sdf.h:20-28 */
static inline float g_mod(float a, float b) { return a - b * floorf(a / b); }     /* GLSL mod */
static float checkboard_pattern(v2 pos, float scale) {                            /* src/util.h:95-101 */
    v2 pattern = V2(floorf(pos.x * scale), floorf(pos.y * scale));
    return g_mod(pattern.x + pattern.y, 2.0f);
}
enum { AO_MAT_DEBUG = 0, AO_MAT_GROUND, AO_MAT_PIPE, AO_MAT_BOTTOM, AO_MAT_DECK, AO_MAT_COPING };   /* :14-19 */
static const v3 ao_size = {1.3f, 1.0f, 1.25f};                                    /* :53 */

static v2 ao_sdf_pipe(v3 pos) {                                                   /* :55-116 */
    const v3 size = ao_size;
    /* ramp: box minus a lying cylinder */
    v3 p = sub3(pos, V3(0, size.y, 0));
    float b = sd_box(p, size);
    p = sub3(p, V3(.7f, .5f, 0));
    p = vec_mat(p, rotate_around_x(-90.0f));
    float c = sd_y_cylinder(p, size.y + .55f, 2.0f * size.z + .1f);
    v2 pipe = V2(op_sub1(b, c), (float)AO_MAT_PIPE);
    /* coping bar */
    p = sub3(pos, V3(0, size.y, 0));
    p = sub3(p, V3(-size.x + .525f, size.y, 0));
    p = vec_mat(p, rotate_around_x(-90.0f));
    v2 coping = V2(sd_y_cylinder(p, .025f, 2.0f * size.z), (float)AO_MAT_COPING);
    /* deck railing */
    p = sub3(pos, V3(0, size.y * 2.0f, 0));
    float rail = sd_box(add3(p, V3(size.x, -.25f, 0)), V3(.025f, .05f, size.z));
    const v3 B = {.025f, .125f, .025f};
    const float H = -.125f;
    float bar_1 = sd_box(add3(p, V3(size.x, H, 0)), B);
    float bar_2 = sd_box(add3(p, V3(size.x, H, size.z / 2.0f)), B);
    float bar_3 = sd_box(add3(p, V3(size.x, H, size.z)), B);
    float bar_4 = sd_box(add3(p, V3(size.x, H, -size.z / 2.0f)), B);
    float bar_5 = sd_box(add3(p, V3(size.x, H, -size.z)), B);
    float b_a = g_min(bar_1, bar_2);
    float b_b = g_min(b_a, bar_3);
    float b_c = g_min(bar_4, bar_5);
    float bars = g_min(b_b, b_c);
    v2 railing = V2(g_min(rail, bars), (float)AO_MAT_DECK);
    v2 deck = op_add2(railing, coping);
    return op_add2(pipe, deck);
}
static v2 ao_sdf(v3 pos) {                                                        /* :118-157 */
    const v3 size = ao_size;
    const float B = .15f;
    v3 p = sub3(pos, V3(0, B, 0));
    v2 bottom = V2(sd_box(p, V3(2.25f * size.x, B, size.z)), (float)AO_MAT_BOTTOM);
    v2 pipe1 = ao_sdf_pipe(add3(p, V3(1.25f * size.x, 0, 0)));
    p = sub3(p, V3(1.25f * size.x, 0, 0));
    p = vec_mat(p, rotate_around_y(180.0f));
    v2 pipe2 = ao_sdf_pipe(p);
    v2 pipe = op_add2(pipe1, pipe2);
    v2 ref = V2(sd_box(pos, V3(.025f, 15.0f, .025f)), (float)AO_MAT_DEBUG);
    v2 ground = V2(sd_plane(pos, V3(0, 1, 0), 0.0f), (float)AO_MAT_GROUND);
    v2 g = op_add2(ground, ref);
    v2 b = op_add2(pipe, bottom);
    return op_add2(b, g);
}
static v3 ao_sdf_normal(v3 p) {                                                   /* :159-170 */
    const float dt = 0.001f;
    v3 x = V3(dt, 0, 0), y = V3(0, dt, 0), z = V3(0, 0, dt);
    return norm3(V3(ao_sdf(add3(p, x)).x - ao_sdf(sub3(p, x)).x,
                    ao_sdf(add3(p, y)).x - ao_sdf(sub3(p, y)).x,
                    ao_sdf(add3(p, z)).x - ao_sdf(sub3(p, z)).x));
}
static float ao_occlusion(const hit_t* hit) {                                     /* sdf_ao, :172-188 (.x of its grey) */
    const float dt = .5f;
    const int steps = 5;
    float occlusion = 0.0f;
    for (float i = 1.0f; i <= (float)steps; i += 1.0f) {
        v3 p = add3(hit->origin, rscale3(dt * i, hit->normal));
        float d = ao_sdf(p).x;
        occlusion += 1.0f / m_pow(2.0f, i) * (dt * i - d);
    }
    return 1.0f - g_clamp(occlusion, 0.0f, 1.0f);
}
static v3 ao_illuminate(v3 eye, const hit_t* hit, float ao, float sh) {           /* :216-250 */
    const v3 sun_dir = norm3(V3(1, 2, 1));                                        /* :214 */
    v3 accum = V3(0, 0, 0);
    float sun_ray = g_max(0.0f, dot3(sun_dir, hit->normal));                      /* key light */
    accum = add3(accum, rscale3(sh * sun_ray, V3(1.2f, 1.3f, 1.0f)));
    float h = hit->normal.y;                                                      /* fill 1: faked hemisphere */
    accum = add3(accum, rscale3(ao * h, V3(.15f, .15f, .4f)));
    float ind = g_max(0.0f, dot3(mul3(sun_dir, V3(-1, 0, -1)), hit->normal));     /* fill 2: indirect */
    accum = add3(accum, rscale3(ao * ind, V3(.4f, .28f, .2f)));
    v3 materials[6];                                                              /* setup_scene, :35-43 */
    materials[AO_MAT_DEBUG] = V3(1, 1, 1);
    materials[AO_MAT_GROUND] = V3(0, .2f, 0);
    materials[AO_MAT_PIPE] = V3(.1f, .1f, .1f);
    materials[AO_MAT_BOTTOM] = materials[AO_MAT_PIPE];
    materials[AO_MAT_DECK] = materials[AO_MAT_PIPE];
    materials[AO_MAT_COPING] = V3(.4f, .4f, .4f);
    v3 mat_c = V3(0, 0, 0);                                                       /* get_material, :22-33: unset outside 0..5 */
    if (hit->material_id >= 0 && hit->material_id < 6) mat_c = materials[hit->material_id];
    if (hit->material_id == AO_MAT_GROUND) {
        float cb = checkboard_pattern(V2(hit->origin.x, hit->origin.z), .5f);
        mat_c = mix3(sub3(mat_c, rscale3(.15f, mat_c)), add3(mat_c, rscale3(.15f, mat_c)), cb);
    }
    (void)eye;                                                                    /* V is computed and unused, :224 */
    return mul3(accum, mat_c);
}
typedef struct { v3 rgb; float w; } v4c;
static v4c ao_render_impl(ray_t ray) {                                            /* :252-294 */
    const int steps = 70;
    const float end = 20.0f;
    float t = 0.0f;
    v4c r;
    for (int i = 0; i < steps; i++) {
        v3 p = add3(ray.origin, scale3(ray.direction, t));
        v2 d = ao_sdf(p);
        if (t > end) break;
        if (d.x < .005f) {
            hit_t h;
            h.t = t;
            h.material_id = (int)d.y;
            h.normal = ao_sdf_normal(p);
            h.origin = p;
            float ao = ao_occlusion(&h);
            float sh = 1.0f;
            r.rgb = ao_illuminate(ray.origin, &h, ao, sh);
            r.w = t;
            return r;
        }
        t += d.x;
    }
    r.rgb = V3(.1f, .1f, .7f);                                                    /* background, :9-12 */
    r.w = t;
    return r;
}
static v3 ao_render(const uniforms_t* u, ray_t ray, v3 point_cam) {               /* :296-320: exponential height fog */
    (void)point_cam;
    v4c orig = ao_render_impl(ray);
    const float t = orig.w;
    const v3 fog_color = V3(1, 1, 1);
    const float density = u->p->fog_density, falloff = u->p->fog_falloff;
    float fog_factor = density * m_exp(-ray.origin.y * falloff) * (1.0f - m_exp(-t * ray.direction.y * falloff)) /
                       (ray.direction.y * falloff);
    return abs3(mix3(orig.rgb, fog_color, fog_factor));
}
static void sdf_ao_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {  /* setup_camera :45-50, FOV 1. */
    m3 rot = rotate_around_y(u->time * 50.0f);
    v3 eye = mat_vec(rot, V3(0, 3, 5));
    main_image(u, ao_render, eye, V3(0, 0, 0), 1.0f, fx, fy, out);
}

/* =================================== APP_VINYL (src/app_vinyl.h) ============================== */
static float sd_capsule(v3 p, v3 a, v3 b, float r) {                              /* src/sdf.h:162-171 */
    v3 ab = sub3(b, a);
    float t = g_clamp(dot3(sub3(p, a), ab) / dot3(ab, ab), 0.0f, 1.0f);
    return len3(sub3(add3(scale3(ab, t), a), p)) - r;
}
static inline m3 M3cols(v3 a, v3 b, v3 c) { m3 m; m.c0 = a; m.c1 = b; m.c2 = c; return m; }
enum { VY_MAT_DEBUG = 0, VY_MAT_GROOVE, VY_MAT_DEAD_WAX, VY_MAT_LABEL, VY_MAT_LOGO, VY_MAT_SHINY };   /* :20-24 */
typedef struct { m3 platter_rot; float time; } vinyl_state;                       /* _mutable platter_rot (:69) + u_time */

static float vy_sdf_logo(v3 pos, float thick) {                                   /* :71-87 */
    v3 b = V3(.25f, thick, 1.2f);
    v3 d = V3(.7f, 0, 0);
    v3 p = vec_mat(pos, rotate_around_y(30.0f));
    float v1 = sd_box(sub3(p, d), b);
    p = vec_mat(pos, rotate_around_y(-30.0f));
    float v2 = sd_box(add3(p, d), b);
    float x = sd_box(pos, V3(1.5f, thick, 1.35f));
    float v = g_min(v1, v2);
    return g_max(v, x);                                                           /* op_intersect, src/sdf.h:30-36 */
}
static v2 vy_sdf_platter(v3 p) {                                                  /* :89-128 */
    const float thick = .1f;
    v2 lead_in = V2(sd_y_cylinder(p, 6.0f, thick - .05f), (float)VY_MAT_DEAD_WAX);
    v2 groove = V2(sd_y_cylinder(p, 5.9f, thick), (float)VY_MAT_GROOVE);
    v2 dead_wax = V2(sd_y_cylinder(p, 3.0f, thick), (float)VY_MAT_DEAD_WAX);
    v2 label = V2(sd_y_cylinder(p, 2.0f, thick), (float)VY_MAT_LABEL);
    v2 logo = V2(vy_sdf_logo(p, thick - .0175f), (float)VY_MAT_LOGO);
    float spc = sd_y_cylinder(p, .10f, .6f);
    float sps = sd_sphere(sub3(p, V3(0, .3f, 0)), .10f);
    v2 spindle = V2(g_min(spc, sps), (float)VY_MAT_SHINY);
    v2 d0 = op_add2(groove, lead_in);
    v2 d1 = op_add2(d0, dead_wax);
    v2 d2 = op_add2(label, logo);
    v2 d3 = op_add2(d1, d2);
    v2 d4 = op_add2(d3, spindle);
    float defect1 = sd_sphere(add3(p, V3(6.05f, 0, 0)), .1f);                     /* notches that make the spin visible */
    float defect2 = sd_sphere(add3(p, V3(-6.05f, 0, 0)), .1f);
    float defect = g_min(defect1, defect2);
    return V2(op_sub1(d4.x, defect), d4.y);
}
static v2 vy_sdf_tonearm(const vinyl_state* st, v3 pos) {                         /* :130-249 */
    v3 base_p = V3(-7, 0, -5);
    float platter = sd_y_cylinder(pos, 6.25f, 1.0f);
    float base_0 = sd_y_cylinder(sub3(pos, base_p), 3.0f, .25f);
    float base_1 = op_sub1(base_0, platter);
    float base_2 = sd_y_cylinder(sub3(pos, base_p), 1.25f, 1.0f);
    float base_12 = g_min(base_1, base_2);
    v2 base_a = V2(base_12, (float)VY_MAT_SHINY);
    v2 base_b = V2(sd_y_cylinder(sub3(pos, base_p), 0.5f, 2.5f), (float)VY_MAT_SHINY);
    v2 base = op_add2(base_a, base_b);

    v3 p = vec_mat(pos, rotate_around_x(m_sin(st->time * 3.6758f) * .1f));        /* needle wobble */

    const float R = .1f, H = .8f;
    v3 a1 = V3(-6, H, -3), a11 = V3(-4.25f, H, 2), a2 = V3(-4.1f, H, 2.45f), a33 = V3(-3.5f, H, 3), a3 = V3(-2, H, 4);
    float arm1 = sd_capsule(p, add3(base_p, V3(-1, H, -2)), a1, R);
    float arm2 = sd_capsule(p, a1, a11, R);
    float arm3 = sd_capsule(p, a33, a3, R);
    v2 armb = sd_bezier(a11, a2, a33, p, R);
    float arm_link1 = g_min(arm1, arm2);
    float arm_link2 = g_min(arm_link1, arm3);
    v2 arm = V2(g_min(arm_link2, armb.x), (float)VY_MAT_SHINY);

    v3 arm_fwd = norm3(sub3(a3, a33));
    v3 arm_up = V3(0, 1, 0);
    v3 arm_right = cross3(arm_fwd, arm_up);
    m3 arm_xform = M3cols(arm_fwd, arm_up, arm_right);

    v3 clr_p = sub3(p, a3);                                                       /* collar */
    float clr_r = R * 1.5f;
    float collar = sd_cylinder(clr_p, V3(0, 0, 0), add3(V3(0, 0, 0), scale3(arm_fwd, .05f)), clr_r);

    const float fl_w = .045f, fl_h = .020f;                                       /* finger lift */
    float fl_len1 = clr_r * 1.0f;
    float fl_len2 = fl_len1 * 1.2f;
    m3 fl_rot = mat_mat(arm_xform, rotate_around_x(45.0f));
    v3 fl_p = vec_mat(sub3(sub3(clr_p, scale3(arm_right, clr_r)), scale3(arm_up, clr_r)), fl_rot);
    float fl1 = sd_box(fl_p, V3(fl_w, fl_h, fl_len1));
    m3 fl_rot2 = rotate_around_x(-45.0f);
    float fl2 = sd_box(sub3(vec_mat(sub3(fl_p, V3(0, 0, fl_len1)), fl_rot2), V3(0, 0, fl_len2)), V3(fl_w, fl_h, fl_len2));
    float finger_lift = g_min(fl1, fl2);
    v2 headshell = V2(g_min(collar, finger_lift), (float)VY_MAT_SHINY);

    const float ctg_w = .05f, ctg_h = .05f;                                       /* cartridge */
    float ctg_len1 = .3f, ctg_len2 = .5f;
    v3 ctg_p = vec_mat(clr_p, arm_xform);
    float ctg1 = sd_box(ctg_p, V3(ctg_len1, ctg_h, ctg_w));
    m3 ctg_rot = rotate_around_z(44.0f);
    v3 ctg2_p = sub3(vec_mat(sub3(ctg_p, V3(ctg_len1, 0, 0)), ctg_rot), V3(ctg_len2 - 0.03f, -.01f, 0));
    float ctg2 = sd_box(ctg2_p, V3(ctg_len2, ctg_h, ctg_w));
    float cut = sd_box(vec_mat(sub3(vec_mat(ctg2_p, rotate_around_x(10.0f)), V3(0, .05f, .175f)), rotate_around_y(-5.0f)),
                       V3(ctg_len2 * 2.0f, ctg_h * 3.0f, ctg_w * 3.2f));
    float cut2 = sd_box(vec_mat(sub3(ctg2_p, V3(.3f, .2f, 0)), rotate_around_z(10.0f)), V3(.4f, .2f, .3f));
    float ctg12 = g_min(ctg1, ctg2);
    float ctg12c = op_sub1(ctg12, cut);
    v2 cartridge = V2(op_sub1(ctg12c, cut2), (float)VY_MAT_SHINY);

    v2 tone1 = op_add2(base, arm);
    v2 tone2 = op_add2(headshell, cartridge);
    return op_add2(tone1, tone2);
}
static v2 vy_sdf(const vinyl_state* st, v3 pos) {                                 /* :251-259 */
    v3 p = vec_mat(pos, st->platter_rot);
    v2 plat = vy_sdf_platter(p);
    v2 arm = vy_sdf_tonearm(st, pos);
    return op_add2(plat, arm);
}
static v3 vy_sdf_normal(const vinyl_state* st, v3 p) {                            /* :261-272 */
    const float dt = 0.001f;
    v3 x = V3(dt, 0, 0), y = V3(0, dt, 0), z = V3(0, 0, dt);
    return norm3(V3(vy_sdf(st, add3(p, x)).x - vy_sdf(st, sub3(p, x)).x,
                    vy_sdf(st, add3(p, y)).x - vy_sdf(st, sub3(p, y)).x,
                    vy_sdf(st, add3(p, z)).x - vy_sdf(st, sub3(p, z)).x));
}
static inline float vy_saw(float x) { return x - floorf(x); }                     /* :274-277 */
static inline float vy_pulse(float x) { return vy_saw(x + .5f) - vy_saw(x); }     /* :279-282 */

static v3 vy_illuminate(const vinyl_state* st, v3 sun_dir, v3 eye, hit_t* hit) {  /* :287-373 */
    v3 L = sun_dir;
    v3 V = norm3(sub3(eye, hit->origin));
    v3 base_color = V3(0, 0, 0);                                                  /* get_material over setup_scene (:40-54) */
    switch (hit->material_id) {
        case VY_MAT_DEBUG: base_color = V3(1, 1, 1); break;
        case VY_MAT_GROOVE: base_color = V3(.01f, .01f, .01f); break;
        case VY_MAT_DEAD_WAX: base_color = V3(.05f, .05f, .05f); break;
        case VY_MAT_LABEL: base_color = V3(.5f, .5f, .0f); break;
        case VY_MAT_LOGO: base_color = V3(0, 0, .7f); break;
        case VY_MAT_SHINY: base_color = V3(.7f, .7f, .7f); break;
        default: break;
    }
    if (hit->material_id == VY_MAT_GROOVE || hit->material_id == VY_MAT_DEAD_WAX) {
        /* Ward anisotropic highlight in the platter's rotating frame */
        hit->origin = vec_mat(hit->origin, st->platter_rot);
        L = vec_mat(L, st->platter_rot);
        V = vec_mat(V, st->platter_rot);
        float r = len3(hit->origin);
        v3 B = divs3(hit->origin, r);
        v3 N = V3(0, 1, 0);
        if (hit->material_id == VY_MAT_GROOVE) {
            float rr = r + .07575f * noise_iq(scale3(hit->origin, 2.456f));
            float s = vy_pulse(rr * 24.0f);
            if (s > 0.0f) {
                N = norm3(add3(N, B));
                N = reflect3(N, V3(0, 1, 0));
            }
        }
        if (hit->material_id == VY_MAT_DEAD_WAX) {
            float s = vy_saw(r * 4.0f);
            N = norm3(add3(N, scale3(B, (float)(s > .9f))));
        }
        v3 T = cross3(B, N);
        const float ro_diff = 1.0f, ro_spec = .0725f, a_x = .025f, a_y = .5f;
        v3 H = norm3(add3(V, L));
        float dotLN = dot3(L, N);
        v3 diffuse = scale3(scale3(base_color, ro_diff / PI_F), g_max(0.0f, dotLN));
        float spec_a = ro_spec / m_sqrt(dotLN * dot3(V, N));
        float spec_b = 1.0f / (4.0f * PI_F * a_x * a_y);
        float ht = dot3(H, T) / a_x;
        float hb = dot3(H, B) / a_y;
        float spec_c = -2.0f * (ht * ht + hb * hb) / (1.0f + dot3(H, N));
        v3 specular = scale3(scale3(scale3(V3(1, 1, 1), spec_a), spec_b), m_exp(spec_c));
        return add3(diffuse, specular);
    }
    hit->normal = vy_sdf_normal(st, hit->origin);
    v3 diffuse = scale3(base_color, g_max(0.0f, dot3(L, hit->normal)));
    v3 H = norm3(add3(V, L));
    v3 specular = rscale3(m_pow(g_max(0.0f, dot3(H, hit->normal)), 50.0f), V3(1, 1, 1));
    return add3(diffuse, specular);
}
static float vy_sdf_shadow(const vinyl_state* st, ray_t ray) {                     /* :375-400 */
    const int steps = 20;
    const float end = 5.0f, penumbra_factor = 16.0f, darkest = .05f;
    float t = 0.0f, umbra = 1.0f;
    for (int i = 0; i < steps; i++) {
        v3 p = add3(ray.origin, scale3(ray.direction, t));
        v2 d = vy_sdf(st, p);
        if (t > end) break;
        if (d.x < .005f) return darkest;
        t += d.x;
        umbra = g_min(umbra, penumbra_factor * d.x / t);
    }
    return umbra;
}
static v3 vy_render(const uniforms_t* u, ray_t ray, v3 point_cam) {               /* :405-455 */
    (void)point_cam;
    const int steps = 60;                                                         /* the __cplusplus branch, :411-416 */
    const float end = 40.0f;
    const v3 sun_dir = norm3(V3(-1, 4, -3));                                      /* :284-285 */
    vinyl_state st;
    st.time = u->time;
    float rot = u->time * 200.0f;
    st.platter_rot = mat_mat(rotate_around_y(rot), rotate_around_x(m_sin(u->time) * .1f));
    float t = 0.0f;
    for (int i = 0; i < steps; i++) {
        v3 p = add3(ray.origin, scale3(ray.direction, t));
        v2 d = vy_sdf(&st, p);
        if (t > end) break;
        if (d.x < .005f) {
            hit_t h;
            h.t = t;
            h.material_id = (int)d.y;
            h.normal = V3(0, 1, 0);
            h.origin = p;
            ray_t sh_ray;
            sh_ray.origin = add3(p, scale3(sun_dir, 0.05f));
            sh_ray.direction = sun_dir;
            float sh = vy_sdf_shadow(&st, sh_ray);
            return scale3(vy_illuminate(&st, sun_dir, ray.origin, &h), sh);
        }
        t += d.x;
    }
    return V3(1, 1, 1);                                                           /* background, :15-18 */
}
static void vinyl_pixel(const uniforms_t* u, float fx, float fy, float out[4]) {  /* setup_camera :56-67, FOV 1. */
    main_image(u, vy_render, V3(0, 5.75f, 6.75f), V3(0, -2.5f, 0), 1.0f, fx, fy, out);
}

/* ============ the 3-D noise volume of util/ddsvolgen (src/noise_worley.h + src/fbm.h:8) ========= */
static inline v3 fract3(v3 a) { return V3(g_fract(a.x), g_fract(a.y), g_fract(a.z)); }
static inline v3 floor3(v3 a) { return V3(floorf(a.x), floorf(a.y), floorf(a.z)); }
static v3 hash_w(v3 x) {                                                          /* src/noise_worley.h:5-17 */
    v3 xx = V3(dot3(x, V3(127.1f, 311.7f, 74.7f)), dot3(x, V3(269.5f, 183.3f, 246.1f)), dot3(x, V3(113.5f, 271.9f, 124.6f)));
    return fract3(V3(m_sin(xx.x) * 43758.5453123f, m_sin(xx.y) * 43758.5453123f, m_sin(xx.z) * 43758.5453123f));
}
static v3 noise_w(v3 pos, float domain_repeat) {                                  /* src/noise_worley.h:20-51 */
    v3 x = scale3(pos, domain_repeat);
    v3 p = floor3(x), f = fract3(x);
    float id = 0.0f;
    v2 res = V2(100.0f, 100.0f);
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                v3 b = V3((float)i, (float)j, (float)k);
                v3 pb = add3(p, b);
                v3 cell = V3(g_mod(pb.x, domain_repeat), g_mod(pb.y, domain_repeat), g_mod(pb.z, domain_repeat));
                v3 r = add3(sub3(b, f), hash_w(cell));
                float d = dot3(r, r);
                if (d < res.x) {
                    id = dot3(add3(p, b), V3(1.0f, 57.0f, 113.0f));
                    res = V2(d, res.x);
                } else if (d < res.y) {
                    res.y = d;
                }
            }
    return V3(m_sqrt(res.x), m_sqrt(res.y), fabsf(id));
}
/* DECL_FBM_FUNC_TILE(fbm_worley_tile, 4, (1. - (noise_w(p, L).r + .25))) and fbm_dds (util/ddsvolgen/src/ddsvolgen.cpp:52-61) */
static float fbm_dds(v3 pos) {
    const float lacunarity = 2.0f, init_gain = 1.0f, gain = .5f;
    float H = init_gain, L = lacunarity, t = 0.0f;
    for (int i = 0; i < 4; i++) {
        t += (1.0f - (noise_w(pos, L).x + .25f)) * H;
        L *= lacunarity;
        H *= gain;
    }
    return t;
}
/* slices [z0, z0 + nz) of the size^3 RGBA32F volume, x fastest (ddsvolgen.cpp:101-116) */
int sbxoracle_bake_volume(int size, int z0, int nz, float* out) {
    if (!out || size <= 0 || z0 < 0 || nz < 0 || z0 + nz > size) return SBX_ERR_INVALID;
    float* ptr = out;
    for (int z = z0; z < z0 + nz; z++)
        for (int y = 0; y < size; y++)
            for (int x = 0; x < size; x++) {
                v3 pos = divs3(adds3(V3((float)x, (float)y, (float)z), .5f), (float)size);
                *ptr++ = fbm_dds(pos);
                *ptr++ = 0.0f;
                *ptr++ = 0.0f;
                *ptr++ = 0.0f;
            }
    return SBX_OK;
}

/* ======================================= frame driver ========================================= */
typedef void (*pixel_fn)(const uniforms_t*, float, float, float[4]);
typedef struct {
    pixel_fn fn;
    uniforms_t u;
    const int* rows;
    int n_rows, width;
    float* out;
    int* next;
    pthread_mutex_t* lock;
    counts_t counts;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    memset(&tl, 0, sizeof tl);
    for (;;) {
        pthread_mutex_lock(j->lock);
        int k = (*j->next)++;
        pthread_mutex_unlock(j->lock);
        if (k >= j->n_rows) break;
        const int y = j->rows[k];
        float* dst = j->out + (size_t)k * (size_t)j->width * 4;
        for (int x = 0; x < j->width; ++x) j->fn(&j->u, (float)x + 0.5f, (float)y + 0.5f, dst + 4 * x);
    }
    j->counts = tl;
    return 0;
}

static int render_frame(pixel_fn fn, const sbx_params* p, const sbx_shard* shard, float* out, int nthreads,
                        unsigned long long* counts_out) {
    if (!p || !out || p->width <= 0 || p->height <= 0) return SBX_ERR_INVALID;
    sbx_shard s = {1, 1, 0};
    if (shard && shard->n_parts > 0) s = *shard;
    if (s.stripe_rows <= 0) s.stripe_rows = 1;
    if (s.part < 0 || s.part >= s.n_parts) return SBX_ERR_INVALID;
    int* rows = (int*)malloc(sizeof(int) * (size_t)p->height);
    int n_rows = 0;
    for (int y = 0; y < p->height; ++y)
        if ((y / s.stripe_rows) % s.n_parts == s.part) rows[n_rows++] = y;
    if (nthreads <= 0) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    int next = 0;
    job_t* jobs = (job_t*)calloc((size_t)nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        job_t* j = &jobs[t];
        j->fn = fn;
        j->u.p = p;
        j->u.res = V2((float)p->width, (float)p->height);
        j->u.time = p->u_time;
        memcpy(j->u.mouse, p->u_mouse, sizeof j->u.mouse);
        j->rows = rows; j->n_rows = n_rows; j->width = p->width; j->out = out;
        j->next = &next; j->lock = &lock;
        if (nthreads == 1) worker(j);
        else pthread_create(&th[t], 0, worker, j);
    }
    if (nthreads > 1)
        for (int t = 0; t < nthreads; ++t) pthread_join(th[t], 0);
    if (counts_out) {
        memset(counts_out, 0, 7 * sizeof(unsigned long long));
        for (int t = 0; t < nthreads; ++t) {
            counts_out[0] += jobs[t].counts.n_sin; counts_out[1] += jobs[t].counts.n_cos;
            counts_out[2] += jobs[t].counts.n_exp; counts_out[3] += jobs[t].counts.n_pow;
            counts_out[4] += jobs[t].counts.n_sqrt; counts_out[5] += jobs[t].counts.n_other;
        }
    }
    free(jobs); free(th); free(rows);
    return SBX_OK;
}

#define ORACLE_ENTRY(name, fn)                                                                       \
    int sbxoracle_render_##name(const sbx_params* p, const sbx_shard* shard, float* out, int nthreads, \
                                unsigned long long* counts_out) {                                    \
        return render_frame(fn, p, shard, out, nthreads, counts_out);                                \
    }
ORACLE_ENTRY(egg, egg_pixel)
ORACLE_ENTRY(clouds, clouds_pixel)
ORACLE_ENTRY(clouds_tex, clouds_tex_pixel)
ORACLE_ENTRY(atmosphere, atmosphere_pixel)
ORACLE_ENTRY(planet, planet_pixel)
ORACLE_ENTRY(raytracer, raytracer_pixel)
ORACLE_ENTRY(sdf_ao, sdf_ao_pixel)
ORACLE_ENTRY(vinyl, vinyl_pixel)
