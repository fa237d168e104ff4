// def.h -- device binding of the shaderbox language layer (replaces src/def.h:1-84 for the CUDA
// build).  Included by an app header INSIDE the per-pixel `sbx_app` object (sbx_kernel.cuh), so
// everything declared here is a member: "_mutable" state is per pixel by construction, which is
// the GLSL meaning of the reference's file-scope variables (src/def.h:16-17).
//
// Parameter-passing macros match the reference's C++ branch (src/def.h:2-9).
#define _in(T) const T&
#define _inout(T) T&
#define _out(T) T&
#define _begin(type) type {
#define _end }
#define _mutable(T) T
#define _constant(T) const T
#define mul(a, b) (a) * (b)

#include "uniform_buffer.h"

#define PI 3.14159265359f

// src/def.h:53-56
struct ray_t {
    vec3 origin;
    vec3 direction;
};
#define BIAS 1e-4f

// src/def.h:59-69
struct sphere_t {
    vec3 origin;
    float radius;
    int material;
};
struct plane_t {
    vec3 direction;
    float distance;
    int material;
};

// src/def.h:71-83 -- a miss is t = 1e8 + 10 and material -1
struct hit_t {
    float t;
    int material_id;
    vec3 normal;
    vec3 origin;
};
#define max_dist 1e8f
const hit_t no_hit = hit_t{float(max_dist + 1e1f), -1, vec3(0.0f, 0.0f, 0.0f), vec3(0.0f, 0.0f, 0.0f)};

// HLSL resource types of the USE_NOISE_TEX branch (src/app_clouds.h:51-55), for app headers compiled with that define
#if defined(USE_NOISE_TEX) && !defined(SBX_NATIVE_NOISE_TEX)
#include "hlsl_tex.h"
#endif
