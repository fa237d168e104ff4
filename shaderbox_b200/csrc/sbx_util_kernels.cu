// sbx_util_kernels.cu -- small helper kernels of the host library (compiled to sbx_util.cubin):
//   sbx_hash_table_kernel  fills the lattice-hash memo used by noise_iq.h
//   sbx_unshard_kernel     scatters a compacted row shard into the full frame (after the gather)
//   sbx_eval_op_kernel     evaluates one operator of the device library on n inputs (test hook)
// Built with the same strict flags as the render kernels (--fmad=false ...).
#include "sbx/sbx_launch.h"
#include "sbx/sbx_vec.cuh"

__device__ __forceinline__ void sbx_util_stage_lut(const void* lut_global) {
    // plain cooperative copy: these kernels are not hot
    const unsigned long long* src = (const unsigned long long*)lut_global;
    unsigned long long* dst = (unsigned long long*)sbx_smem;
    for (int i = threadIdx.x; i < SBX_LUT_MATH_BYTES / 8; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// hash(n) = fract(sin(n) * 753.5453123)   (src/noise_iq.h:5-9)
__device__ __forceinline__ float sbx_hash_of(float n) {
    const float s = sbx_sinf(n) * 753.5453123f;
    return s - floorf(s);
}
// entry k = the eight corners of the noise_iq cell with base index n = lo + k, z-neighbours adjacent
// (layout and rationale: include/sbx/noise_iq.h)
extern "C" __global__ void sbx_hash_table_kernel(float4* tab, int lo, int len, const void* lut) {
    sbx_util_stage_lut(lut);
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= len) return;
    const float n = (float)(lo + k);
    tab[2 * k] = make_float4(sbx_hash_of(n), sbx_hash_of(n + 113.0f), sbx_hash_of(n + 1.0f), sbx_hash_of(n + 114.0f));
    tab[2 * k + 1] = make_float4(sbx_hash_of(n + 157.0f), sbx_hash_of(n + 270.0f), sbx_hash_of(n + 158.0f), sbx_hash_of(n + 271.0f));
}

// part -> frame: local row lr of shard (stripe, parts, part) is frame row y
extern "C" __global__ void sbx_unshard_kernel(const float4* __restrict__ part, float4* __restrict__ frame,
                                              int width, int local_rows, int stripe, int parts, int which) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)width * local_rows;
    if (i >= total) return;
    const int lr = (int)(i / width), x = (int)(i % width);
    const int y = ((lr / stripe) * parts + which) * stripe + lr % stripe;
    frame[(long long)y * width + x] = part[i];
}

namespace sbx_glsl {
// the operator library, instantiated outside any app: a minimal host object supplies sbx_L
struct sbx_ops {
    const sbx_launch* __restrict__ sbx_L;
    vec2 iResolution;
    float iGlobalTime;
    vec4 iMouse;
    SBX_FN float sbx_uniform(const float& v) { return v; }
#include "def.h"
#include "util.h"
#include "util_optics.h"
#include "intersect.h"
#include "sdf.h"
#include "IK.h"
#include "noise_iq.h"
#include "noise_worley.h"
#include "fbm.h"
#define hg_g (.76f)
#include "volumetric.h"
#include "material.h"
#include "light.h"
#include "cornell_box.h"
    DECL_FBM_FUNC(fbm4, 4, noise_iq(p))
    DECL_FBM_FUNC_TILE(fbm_w3, 3, noise_w(p, L).x)
    // util/ddsvolgen/src/ddsvolgen.cpp:52-61: the tiled-Worley fbm baked into the 3-D noise texture
    DECL_FBM_FUNC_TILE(fbm_worley_tile, 4, (1.0f - (noise_w(p, L).r + .25f)))
    SBX_FN float fbm_dds(_in(vec3) pos) { return fbm_worley_tile(pos, 2.0f, 1.0f, .5f); }
    __device__ explicit sbx_ops(const sbx_launch* L) : sbx_L(L), iResolution(1.0f, 1.0f), iGlobalTime(0.0f) {}
};
}  // namespace sbx_glsl

// The 3-D noise texture bake of util/ddsvolgen/src/ddsvolgen.cpp:101-116 (the USE_NOISE_TEX input of
// src/app_clouds.h:51-55): voxel (x, y, z) of a size^3 R32G32B32A32_FLOAT volume, x fastest, holds
// (fbm_dds((vec3(x, y, z) + .5) / size), 0, 0, 0).  One thread per voxel, one float4 store each: the
// reference's four host threads over z-slabs become one grid.  ~330 sines per voxel -- compute bound.
extern "C" __global__ void __launch_bounds__(128)
sbx_bake_volume_kernel(const __grid_constant__ sbx_launch L, float4* __restrict__ out, int size, int z0, int nz) {
    using namespace sbx_glsl;
    sbx_util_stage_lut(L.lut);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)size * size * nz) return;
    const int x = (int)(i % size), y = (int)((i / size) % size), z = z0 + (int)(i / ((long long)size * size));
    sbx_ops ops(&L);
    const vec3 pos = (vec3(float(x), float(y), float(z)) + .5f) / float(size);
    __stcs(out + i, make_float4(ops.fbm_dds(pos), 0.0f, 0.0f, 0.0f));
}

enum {
    OP_SINF = 0, OP_COSF, OP_TANF, OP_EXPF, OP_POWF, OP_ACOSF, OP_ATAN2F, OP_SQRTF, OP_DIVF,
    OP_HASH = 16, OP_HASH_ARITH, OP_NOISE_IQ, OP_NOISE_W, OP_FBM4, OP_FBM_W3,
    OP_SD_SPHERE = 32, OP_SD_BOX, OP_SD_TORUS, OP_SD_Y_CYLINDER, OP_SD_CYLINDER, OP_SD_BEZIER, OP_SD_CAPSULE,
    OP_SD_PLANE, OP_OP_BLEND, OP_IK_SOLVER,
    OP_PHASE_HG = 48, OP_PHASE_RAYLEIGH, OP_PHASE_SCHLICK, OP_PHASE_ISO, OP_FRESNEL, OP_REFLECT, OP_REFRACT,
    OP_COOK_TORRANCE, OP_BLINN_PHONG, OP_INTERSECT_SPHERE, OP_INTERSECT_PLANE,
    OP_ROTATE_X = 64, OP_ROTATE_Y, OP_ROTATE_Z, OP_LINEAR_TO_SRGB, OP_BAND, OP_CHECKBOARD, OP_REMAP,
    OP_PRIMARY_RAY, OP_SMOOTHSTEP, OP_MOD, OP_ORTHO_BASIS, OP_UNORM8
};

// in: n rows of in_stride floats; out: n rows of out_stride floats (unused slots left untouched)
extern "C" __global__ void sbx_eval_op_kernel(const __grid_constant__ sbx_launch L, int op, const float* __restrict__ in,
                                              int in_stride, float* __restrict__ out, int out_stride, int n) {
    using namespace sbx_glsl;
    sbx_util_stage_lut(L.lut);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in + (size_t)i * in_stride;
    float* o = out + (size_t)i * out_stride;
    sbx_ops ops(&L);
#define V3(k) vec3(a[k], a[(k) + 1], a[(k) + 2])
#define OUT3(v) do { const vec3 t_ = (v); o[0] = t_.x; o[1] = t_.y; o[2] = t_.z; } while (0)
    switch (op) {
        case OP_SINF: o[0] = sbx_sinf(a[0]); break;
        case OP_COSF: o[0] = sbx_cosf(a[0]); break;
        case OP_TANF: o[0] = sbx_tanf(a[0]); break;
        case OP_EXPF: o[0] = sbx_expf(a[0]); break;
        case OP_POWF: o[0] = sbx_powf(a[0], a[1]); break;
        case OP_ACOSF: o[0] = sbx_acosf(a[0]); break;
        case OP_ATAN2F: o[0] = sbx_atan2f(a[0], a[1]); break;
        case OP_SQRTF: o[0] = sbx_glsl::sqrt(a[0]); break;
        case OP_DIVF: o[0] = a[0] / a[1]; break;
        case OP_HASH: o[0] = ops.hash(a[0]); break;
        case OP_HASH_ARITH: o[0] = ops.sbx_hash_arith(a[0]); break;
        case OP_NOISE_IQ: o[0] = ops.noise_iq(V3(0)); break;
        case OP_NOISE_W: OUT3(ops.noise_w(V3(0), a[3])); break;
        case OP_FBM4: o[0] = ops.fbm4(V3(0), a[3], a[4], a[5]); break;
        case OP_FBM_W3: o[0] = ops.fbm_w3(V3(0), a[3], a[4], a[5]); break;
        case OP_SD_SPHERE: o[0] = ops.sd_sphere(V3(0), a[3]); break;
        case OP_SD_BOX: o[0] = ops.sd_box(V3(0), V3(3)); break;
        case OP_SD_TORUS: o[0] = ops.sd_torus(V3(0), a[3], a[4]); break;
        case OP_SD_Y_CYLINDER: o[0] = ops.sd_y_cylinder(V3(0), a[3], a[4]); break;
        case OP_SD_CYLINDER: o[0] = ops.sd_cylinder(V3(0), V3(3), V3(6), a[9]); break;
        case OP_SD_BEZIER: { const vec2 r = ops.sd_bezier(V3(0), V3(3), V3(6), V3(9), a[12]); o[0] = r.x; o[1] = r.y; break; }
        case OP_SD_CAPSULE: o[0] = ops.sd_capsule(V3(0), V3(3), V3(6), a[9]); break;
        case OP_SD_PLANE: o[0] = ops.sd_plane(V3(0), V3(3), a[6]); break;
        case OP_OP_BLEND: o[0] = ops.op_blend(a[0], a[1], a[2]); break;
        case OP_IK_SOLVER: OUT3(ops.ik_solver(V3(0), V3(3), a[6], a[7])); break;
        case OP_PHASE_HG: o[0] = ops.henyey_greenstein_phase_func(a[0]); break;
        case OP_PHASE_RAYLEIGH: o[0] = ops.rayleigh_phase_func(a[0]); break;
        case OP_PHASE_SCHLICK: o[0] = ops.schlick_phase_func(a[0]); break;
        case OP_PHASE_ISO: o[0] = ops.isotropic_phase_func(a[0]); break;
        case OP_FRESNEL: o[0] = ops.fresnel_factor(a[0], a[1], a[2]); break;
        case OP_REFLECT: OUT3(ops.reflect(V3(0), V3(3))); break;
        case OP_REFRACT: OUT3(ops.refract(V3(0), V3(3), a[6])); break;
        case OP_COOK_TORRANCE:
        case OP_BLINN_PHONG: {
            sbx_ops::hit_t h; h.t = 1.0f; h.material_id = 1; h.normal = V3(6); h.origin = vec3(0.0f, 0.0f, 0.0f);
            sbx_ops::material_t m; m.base_color = V3(9); m.metallic = 0.0f; m.roughness = a[12]; m.ior = a[13];
            m.reflectivity = 0.0f; m.translucency = 0.0f;
            if (op == OP_COOK_TORRANCE) OUT3(ops.illum_cook_torrance(V3(0), V3(3), h, m));
            else OUT3(ops.illum_blinn_phong(V3(0), V3(3), h, m));
            break;
        }
        case OP_INTERSECT_SPHERE: {
            sbx_ops::ray_t r; r.origin = V3(0); r.direction = V3(3);
            sbx_ops::sphere_t s; s.origin = V3(6); s.radius = a[9]; s.material = 3;
            sbx_ops::hit_t h = ops.no_hit;
            ops.intersect_sphere(r, s, h);
            o[0] = h.t; o[1] = (float)h.material_id; o[2] = h.normal.x; o[3] = h.normal.y; o[4] = h.normal.z;
            o[5] = h.origin.x; o[6] = h.origin.y; o[7] = h.origin.z;
            break;
        }
        case OP_INTERSECT_PLANE: {
            sbx_ops::ray_t r; r.origin = V3(0); r.direction = V3(3);
            sbx_ops::plane_t p; p.direction = V3(6); p.distance = a[9]; p.material = 2;
            sbx_ops::hit_t h = ops.no_hit;
            ops.intersect_plane(r, p, h);
            o[0] = h.t; o[1] = (float)h.material_id; o[2] = h.normal.x; o[3] = h.normal.y; o[4] = h.normal.z;
            o[5] = h.origin.x; o[6] = h.origin.y; o[7] = h.origin.z;
            break;
        }
        case OP_ROTATE_X: OUT3(ops.rotate_around_x(a[0]) * V3(1)); break;
        case OP_ROTATE_Y: OUT3(ops.rotate_around_y(a[0]) * V3(1)); break;
        case OP_ROTATE_Z: OUT3(V3(1) * ops.rotate_around_z(a[0])); break;
        case OP_LINEAR_TO_SRGB: OUT3(ops.linear_to_srgb(V3(0))); break;
        case OP_BAND: o[0] = ops.band(a[0], a[1], a[2], a[3]); break;
        case OP_CHECKBOARD: o[0] = ops.checkboard_pattern(vec2(a[0], a[1]), a[2]); break;
        case OP_REMAP: o[0] = ops.remap(a[0], a[1], a[2], a[3], a[4]); break;
        case OP_PRIMARY_RAY: {
            vec3 eye = V3(3), look = V3(6);
            const sbx_ops::ray_t r = ops.get_primary_ray(V3(0), eye, look);
            o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
            o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z;
            break;
        }
        case OP_SMOOTHSTEP: o[0] = smoothstep(a[0], a[1], a[2]); break;
        case OP_MOD: o[0] = mod(a[0], a[1]); break;
        case OP_ORTHO_BASIS: { vec3 f, r; ops.fast_orthonormal_basis(V3(0), f, r); o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = r.x; o[4] = r.y; o[5] = r.z; break; }
        case OP_UNORM8: o[0] = (float)sbx_unorm8(a[0]); break;
        default: break;
    }
#undef V3
#undef OUT3
}
