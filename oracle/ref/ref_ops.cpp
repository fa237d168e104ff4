// ref_ops.cpp -- TEST INFRASTRUCTURE (oracle/_ref).  Never linked into the product library.
//
// Operator known-answer side of the parity tests: the reference's OWN operator headers
// (src/util.h, intersect.h, sdf.h, IK.h, noise_iq.h, noise_worley.h, fbm.h, volumetric.h,
// material.h, light.h, util_optics.h, cornell_box.h -- included verbatim from where they lie under
// /root/reference/src, never copied) evaluated on n inputs.  The op names, input layouts and
// output layouts are those of sbx_eval_op (include/sbx.h), so a test feeds the same array to both.
#include <cstring>

#include "../../include/sbx.h"
#include "glsl_shim.h"

namespace ref_ops {
using namespace glsl;

#define thread_local /* def.h:7-8 -> plain per-instance members */
struct ops_t {
    vec2 iResolution;
    float iGlobalTime;
    vec4 iMouse;
#include "def.h"
#include "util.h"
#include "util_optics.h"
#include "intersect.h"
#include "sdf.h"
#include "IK.h"
#include "noise_iq.h"
#include "noise_worley.h"
#include "fbm.h"
#define hg_g (.76)
#include "volumetric.h"
#include "material.h"
#include "light.h"
#include "cornell_box.h"
    DECL_FBM_FUNC(fbm4, 4, noise_iq(p))
    DECL_FBM_FUNC_TILE(fbm_w3, 3, noise_w(p, L).x)
    // util/ddsvolgen/src/ddsvolgen.cpp:52-61, verbatim
    DECL_FBM_FUNC_TILE(fbm_worley_tile, 4, (1. - (noise_w(p, L).r + .25)))
    float fbm_dds(vec3 &pos)
    {
        return fbm_worley_tile(pos, 2., 1., .5);
    }
    ops_t() : iResolution(1.0f, 1.0f), iGlobalTime(0.0f), iMouse(0, 0, 0, 0) {}
};
#undef thread_local

static void eval_one(ops_t& ops, const char* op, const float* a, float* o) {
#define IS(name) (!std::strcmp(op, name))
#define V3(k) vec3(a[k], a[(k) + 1], a[(k) + 2])
#define OUT3(v) do { const vec3 t_ = (v); o[0] = t_.x; o[1] = t_.y; o[2] = t_.z; } while (0)
    if IS("sinf") o[0] = sin(a[0]);
    else if IS("cosf") o[0] = cos(a[0]);
    else if IS("tanf") o[0] = tan(a[0]);
    else if IS("expf") o[0] = exp(a[0]);
    else if IS("powf") o[0] = pow(a[0], a[1]);
    else if IS("acosf") o[0] = acos(a[0]);
    else if IS("atan2f") o[0] = atan(a[0], a[1]);
    else if IS("sqrtf") o[0] = sqrt(a[0]);
    else if IS("divf") o[0] = a[0] / a[1];
    else if (IS("hash") || IS("hash_arith")) o[0] = ops.hash(a[0]);
    else if IS("noise_iq") o[0] = ops.noise_iq(V3(0));
    else if IS("noise_w") OUT3(ops.noise_w(V3(0), a[3]));
    else if IS("fbm4") o[0] = ops.fbm4(V3(0), a[3], a[4], a[5]);
    else if IS("fbm_w3") o[0] = ops.fbm_w3(V3(0), a[3], a[4], a[5]);
    else if IS("sd_sphere") o[0] = ops.sd_sphere(V3(0), a[3]);
    else if IS("sd_box") o[0] = ops.sd_box(V3(0), V3(3));
    else if IS("sd_torus") o[0] = ops.sd_torus(V3(0), a[3], a[4]);
    else if IS("sd_y_cylinder") o[0] = ops.sd_y_cylinder(V3(0), a[3], a[4]);
    else if IS("sd_cylinder") o[0] = ops.sd_cylinder(V3(0), V3(3), V3(6), a[9]);
    else if IS("sd_bezier") { const vec2 r = ops.sd_bezier(V3(0), V3(3), V3(6), V3(9), a[12]); o[0] = r.x; o[1] = r.y; }
    else if IS("sd_capsule") o[0] = ops.sd_capsule(V3(0), V3(3), V3(6), a[9]);
    else if IS("sd_plane") o[0] = ops.sd_plane(V3(0), V3(3), a[6]);
    else if IS("op_blend") o[0] = ops.op_blend(a[0], a[1], a[2]);
    else if IS("ik_solver") OUT3(ops.ik_solver(V3(0), V3(3), a[6], a[7]));
    else if IS("henyey_greenstein_phase_func") o[0] = ops.henyey_greenstein_phase_func(a[0]);
    else if IS("rayleigh_phase_func") o[0] = ops.rayleigh_phase_func(a[0]);
    else if IS("schlick_phase_func") o[0] = ops.schlick_phase_func(a[0]);
    else if IS("isotropic_phase_func") o[0] = ops.isotropic_phase_func(a[0]);
    else if IS("fresnel_factor") o[0] = ops.fresnel_factor(a[0], a[1], a[2]);
    else if IS("reflect") OUT3(ops.reflect(V3(0), V3(3)));
    else if IS("refract") OUT3(ops.refract(V3(0), V3(3), a[6]));
    else if (IS("illum_cook_torrance") || IS("illum_blinn_phong")) {
        ops_t::hit_t h; h.t = 1.0f; h.material_id = 1; h.normal = V3(6); h.origin = vec3(0.0f, 0.0f, 0.0f);
        ops_t::material_t m; m.base_color = V3(9); m.metallic = 0.0f; m.roughness = a[12]; m.ior = a[13];
        m.reflectivity = 0.0f; m.translucency = 0.0f;
        if IS("illum_cook_torrance") OUT3(ops.illum_cook_torrance(V3(0), V3(3), h, m));
        else OUT3(ops.illum_blinn_phong(V3(0), V3(3), h, m));
    } else if IS("intersect_sphere") {
        ops_t::ray_t r; r.origin = V3(0); r.direction = V3(3);
        ops_t::sphere_t s; s.origin = V3(6); s.radius = a[9]; s.material = 3;
        ops_t::hit_t h = ops.no_hit;
        ops.intersect_sphere(r, s, h);
        o[0] = h.t; o[1] = (float)h.material_id; o[2] = h.normal.x; o[3] = h.normal.y; o[4] = h.normal.z;
        o[5] = h.origin.x; o[6] = h.origin.y; o[7] = h.origin.z;
    } else if IS("intersect_plane") {
        ops_t::ray_t r; r.origin = V3(0); r.direction = V3(3);
        ops_t::plane_t p; p.direction = V3(6); p.distance = a[9]; p.material = 2;
        ops_t::hit_t h = ops.no_hit;
        ops.intersect_plane(r, p, h);
        o[0] = h.t; o[1] = (float)h.material_id; o[2] = h.normal.x; o[3] = h.normal.y; o[4] = h.normal.z;
        o[5] = h.origin.x; o[6] = h.origin.y; o[7] = h.origin.z;
    }
    else if IS("rotate_around_x") OUT3(ops.rotate_around_x(a[0]) * V3(1));
    else if IS("rotate_around_y") OUT3(ops.rotate_around_y(a[0]) * V3(1));
    else if IS("rotate_around_z") OUT3(V3(1) * ops.rotate_around_z(a[0]));
    else if IS("linear_to_srgb") OUT3(ops.linear_to_srgb(V3(0)));
    else if IS("band") o[0] = ops.band(a[0], a[1], a[2], a[3]);
    else if IS("checkboard_pattern") o[0] = ops.checkboard_pattern(vec2(a[0], a[1]), a[2]);
    else if IS("remap") o[0] = ops.remap(a[0], a[1], a[2], a[3], a[4]);
    else if IS("get_primary_ray") {
        vec3 eye = V3(3), look = V3(6);
        const ops_t::ray_t r = ops.get_primary_ray(V3(0), eye, look);
        o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
        o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z;
    }
    else if IS("smoothstep") o[0] = smoothstep(a[0], a[1], a[2]);
    else if IS("mod") o[0] = mod(a[0], a[1]);
    else if IS("fast_orthonormal_basis") {
        vec3 f, r; ops.fast_orthonormal_basis(V3(0), f, r);
        o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = r.x; o[4] = r.y; o[5] = r.z;
    }
#undef IS
#undef V3
#undef OUT3
}

}  // namespace ref_ops

extern "C" int sbxref_eval_op(const char* op, const float* in, int in_stride, float* out, int out_stride, int n) {
    if (!op || !in || !out || in_stride <= 0 || out_stride <= 0 || n < 0) return SBX_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        ref_ops::ops_t ops;   // fresh per-invocation state, as in ref_app.cpp
        ref_ops::eval_one(ops, op, in + (size_t)i * in_stride, out + (size_t)i * out_stride);
    }
    return SBX_OK;
}


// ---- util/ddsvolgen/src/ddsvolgen.cpp restated around the reference's own headers ------------------
// The tool itself cannot be built here (VML + the Windows SDK); its per-voxel expression (:108-109) and
// its header initialisation (:69-92) are reproduced with the reference's noise_worley.h / fbm.h (above)
// and the reference's vendored DDS.h.
#include <cstdint>
typedef uint32_t DWORD;   // ddsvolgen.cpp:9 `typedef unsigned long DWORD` is 32 bits on its platform (Windows)
#include "DDS.h"
namespace ref_dds {
using namespace DirectX;
struct DDS {   // ddsvolgen.cpp:13-18
    DWORD dwMagic;
    DDS_HEADER header;
    DDS_HEADER_DXT10 header10;
};
}

extern "C" {

// slices [z0, z0+nz) of the size^3 volume, RGBA32F, x fastest (ddsvolgen.cpp:101-116)
int sbxref_bake_volume(int size, int z0, int nz, float* out) {
    if (!out || size <= 0 || z0 < 0 || nz < 0 || z0 + nz > size) return SBX_ERR_INVALID;
    using namespace ref_ops;
    ops_t ops;
    float* ptr = out;
    for (size_t z = (size_t)z0; z < (size_t)(z0 + nz); z++)
        for (size_t y = 0; y < (size_t)size; y++)
            for (size_t x = 0; x < (size_t)size; x++) {
                vec3 pos = (vec3(x, y, z) + .5f) / float(size);
                *ptr++ = ops.fbm_dds(pos);
                *ptr++ = 0.f;
                *ptr++ = 0.f;
                *ptr++ = 0.f;
            }
    return SBX_OK;
}

// the bytes ddsvolgen writes before the data (fwrite(&dds, sizeof(dds), 1, ...), :146) for a size^3 volume
int sbxref_dds_header(int size_, unsigned char* out, int capacity) {
    using namespace ref_dds;
    const size_t size = (size_t)size_;
    const size_t channels = 4;
    using FLOAT = float;
    DDS dds = { 0 };
    dds.dwMagic = DDS_MAGIC;
    dds.header.dwSize = sizeof(DDS_HEADER);
    dds.header.dwFlags = DDS_HEADER_FLAGS_TEXTURE | DDS_HEADER_FLAGS_VOLUME | DDS_HEADER_FLAGS_PITCH;
    dds.header.dwHeight = size;
    dds.header.dwWidth = size;
    dds.header.dwDepth = size;
    dds.header.dwPitchOrLinearSize = (size * (sizeof(FLOAT) * channels) + 7) / 8;
    dds.header.dwMipMapCount = 0;
    dds.header.ddspf = DDSPF_DX10;
    dds.header.dwCaps = DDS_SURFACE_FLAGS_TEXTURE | DDS_SURFACE_FLAGS_CUBEMAP;
    dds.header.dwCaps2 = DDS_FLAGS_VOLUME;
    DXGI_FORMAT fmt[] = { DXGI_FORMAT_R32_FLOAT, DXGI_FORMAT_R32G32_FLOAT, DXGI_FORMAT_R32G32B32_UINT, DXGI_FORMAT_R32G32B32A32_FLOAT };
    dds.header10.dxgiFormat = fmt[channels - 1];
    dds.header10.resourceDimension = DDS_DIMENSION_TEXTURE3D;
    dds.header10.arraySize = 1;
    dds.header10.miscFlag = 0;
    dds.header10.miscFlags2 = 0;
    if (!out || capacity < (int)sizeof(dds)) return SBX_ERR_INVALID;
    std::memcpy(out, &dds, sizeof(dds));
    return (int)sizeof(dds);
}

}  // extern "C"
