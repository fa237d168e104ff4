// app_clouds_tex_native.h -- APP_CLOUDS with USE_NOISE_TEX (src/app_clouds.h:8-9): the branch of the cloud app whose
// density comes from two 3-D noise textures instead of the fbm (:51-55, :69-81).  In the reference it exists for the
// HLSL hosts only (`Texture3D.SampleLevel`, sampler state util/hlsltoy/src/hlsltoy.cpp:244-249: MIN_MAG_MIP_LINEAR,
// WRAP on u, v, w; textures baked by util/ddsvolgen); here it is the app "APP_CLOUDS_TEX" with the sampler DEFINED in
// software -- the rule is written down in oracle/sbx_oracle.c (tex_coord / tex_sample_r; D3D11 linear filtering with
// 8-bit sub-texel weights) and this file follows it operation for operation.  PARITY UNPINNED against real hardware
// filtering; bit-identical to the oracle.
//
// The textures live in HBM as padded single-channel volumes (sbx_tex_params, sbx_launch.h).  A warp's 32 rays sample
// points less than a texel apart (a texel is 7.8 world units; neighbouring rays are ~0.4 apart, a march step ~1), so
// the 2x2x2 texel neighbourhoods of all its lanes almost always fit one 8x4x4 box of texels -- and stay inside it for
// several march steps and for the whole 6-step light march.  The warp therefore STAGES that box of both textures into
// shared memory with TMA (two cp.async.bulk.tensor.3d loads completing on the warp's mbarrier, UTMALDG in the SASS)
// whenever the lanes' common box moves, and every lane takes its 16 texels from shared memory.  A sample whose lanes do
// not fit one box (rays near the horizon, where neighbouring pixels are many texels apart) reads global memory instead.
#include "def.h"
#include "util.h"
#include "intersect.h"

#define hg_g (.2f)
#include "volumetric.h"

#define cld_noise_factor .001f
#define SBX_TEX_BOX 4                         // texels of the staged box along y and z
#define SBX_TEX_BOX_X 8                       // ... and along x: TMA wants the box to START on a 16-byte boundary of the row
                                              // (an innermost coordinate that is not a multiple of 4 texels faults with "illegal
                                              // instruction": tools/ubench/tma3d_b.cu), so the box is 8 wide and starts at x & ~3
#define SBX_TEX_BOX_FLOATS (SBX_TEX_BOX_X * SBX_TEX_BOX * SBX_TEX_BOX)

// ---- per-warp staging state (shared memory: the warp's lanes are not always all active -- the light march runs in
// the lanes whose sample met cloud -- so nothing about the staged box may live in a lane's registers) -----------------
struct __align__(128) sbx_tex_stage {       // TMA writes the tiles: 128-byte aligned, one slot per warp
    float tile[2][SBX_TEX_BOX_FLOATS];         // the box of texture 0 and of texture 1 (z, y, x: x fastest), 2 x 512 bytes
    unsigned long long bar;                   // mbarrier: both loads of a refill complete on it
    int bx, by, bz;                           // origin (padded texel coordinates) of the staged box; bx < 0: none yet
    unsigned parity;                          // parity the next completion of `bar` will have
};
SBX_FN sbx_tex_stage* sbx_stage() {
    __shared__ __align__(128) sbx_tex_stage stages[SBX_WARPS_PER_CTA];
    return &stages[threadIdx.x >> 5];
}

// the sampler rule itself (sbx_tex_axis / sbx_tex_lerp / sbx_tex_blend) is the library's: hlsl_tex.h
#include "hlsl_tex.h"

// both textures at pos (already scaled by cld_noise_factor): .x = u_tex_noise.r, .y = u_tex_noise_2.r   (:69, :77).
// WARP-COLLECTIVE: every lane of sbx_lanes calls it, in lock-step (the march loops below are written warp-uniform for
// that reason -- a warp that has split into independently scheduled groups would have two groups refilling the one
// staging slot and flipping the one barrier at the same time); `active` says whether this lane wants the sample.
SBX_FN vec2 sbx_sample_both(_in(vec3) pos, bool active) {
    const sbx_tex_params* T = sbx_T;
    const unsigned mask = sbx_lanes;
    const float n = float(T->size);
    int ix = 0, iy = 0, iz = 0;
    float wx = 0.0f, wy = 0.0f, wz = 0.0f;
    if (active) {
        sbx_tex_axis(pos.x, n, ix, wx);
        sbx_tex_axis(pos.y, n, iy, wy);
        sbx_tex_axis(pos.z, n, iz, wz);
    }
    __syncwarp(mask);
    if (!__any_sync(mask, active)) return vec2(0.0f, 0.0f);
    // the box that would hold every active lane's neighbourhood: origin = the lanes' smallest indices (x rounded down to a
    // multiple of 4 texels = 16 bytes, which TMA requires of the innermost coordinate)
    const int big = 0x7fffffff;
    const int bx = __reduce_min_sync(mask, active ? ix : big) & ~3, by = __reduce_min_sync(mask, active ? iy : big),
              bz = __reduce_min_sync(mask, active ? iz : big);
    const int ex = __reduce_max_sync(mask, active ? ix : -1) - bx, ey = __reduce_max_sync(mask, active ? iy : -1) - by,
              ez = __reduce_max_sync(mask, active ? iz : -1) - bz;
    float c0[8], c1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) c0[k] = c1[k] = 0.0f;
    if (ex <= SBX_TEX_BOX_X - 2 && ey <= SBX_TEX_BOX - 2 && ez <= SBX_TEX_BOX - 2) {      // warp-uniform
        sbx_tex_stage* S = sbx_stage();
        volatile sbx_tex_stage* Sv = S;
        int sx = Sv->bx, sy = Sv->by, sz = Sv->bz;
        // reuse the staged box if every lane's neighbourhood is still inside it, else fetch the box at the lanes' minimum
        const bool inside = sx >= 0 && bx >= sx && by >= sy && bz >= sz && bx + ex <= sx + SBX_TEX_BOX_X - 2 &&
                            by + ey <= sy + SBX_TEX_BOX - 2 && bz + ez <= sz + SBX_TEX_BOX - 2;
        if (!inside) {                                          // warp-uniform
            const unsigned parity = Sv->parity;
            const bool leader = (threadIdx.x & 31) == (__ffs(mask) - 1);
            __syncwarp(mask);                                   // every lane has read the state and is done with the old box
            if (leader) {
                const unsigned bar = sbx_smem_addr(&S->bar);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(2 * SBX_TEX_BOX_FLOATS * 4) : "memory");
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    asm volatile(
                        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                        ::"r"(sbx_smem_addr(S->tile[t])), "l"(reinterpret_cast<const void*>(T->map[t])), "r"(bx), "r"(by), "r"(bz), "r"(bar)
                        : "memory");
                Sv->bx = bx; Sv->by = by; Sv->bz = bz;
                Sv->parity = parity ^ 1u;
            }
            unsigned done = 0;
            while (!done)
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(done) : "r"(sbx_smem_addr(&S->bar)), "r"(parity) : "memory");
            __syncwarp(mask);                                   // the leader's state update is visible to every lane
            sx = bx; sy = by; sz = bz;
        }
        if (active) {
            const int at = ((iz - sz) * SBX_TEX_BOX + (iy - sy)) * SBX_TEX_BOX_X + (ix - sx);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int o = at + (k & 1) + ((k >> 1) & 1) * SBX_TEX_BOX_X + (k >> 2) * SBX_TEX_BOX_X * SBX_TEX_BOX;
                c0[k] = Sv->tile[0][o];
                c1[k] = Sv->tile[1][o];
            }
        }
    } else if (active) {
        const size_t at = (size_t)iz * T->pitch_xy + (size_t)iy * T->pitch_x + ix;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const size_t o = at + (k & 1) + (size_t)((k >> 1) & 1) * T->pitch_x + (size_t)(k >> 2) * T->pitch_xy;
            c0[k] = __ldg(T->vol[0] + o);
            c1[k] = __ldg(T->vol[1] + o);
        }
    }
    return vec2(sbx_tex_blend(c0, wx, wy, wz), sbx_tex_blend(c1, wx, wy, wz));
}

SBX_FN void setup_camera(_inout(vec3) eye, _inout(vec3) look_at) {   // :23-30
    eye = vec3(0.0f, -.5f, 0.0f);
    const float angle = u_mouse.x * .5f;
    look_at = mul(rotate_around_y(angle), vec3(0.0f, 0.0f, -1.0f));
}

SBX_FN void setup_scene() {
    // the warp's staging slot: no box yet, barrier armed, first completion has parity 0 (one lane initialises)
    const unsigned mask = sbx_lanes;
    __syncwarp(mask);
    if ((threadIdx.x & 31) == (__ffs(mask) - 1)) {
        sbx_tex_stage* S = sbx_stage();
        S->bx = S->by = S->bz = -1;
        S->parity = 0u;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbx_smem_addr(&S->bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the TMA unit (async proxy) sees the initialised barrier
    }
    __syncwarp(mask);
}

SBX_FN vec3 render_sky_color(_in(vec3) eye_dir) {   // :36-46
    const float sun_amount = max(dot(eye_dir, sun_dir), 0.0f);
    vec3 sky = mix(vec3(.0f, .1f, .4f), vec3(.3f, .6f, .8f), 1.0f - eye_dir.y);
    sky += sun_color * min(pow(sun_amount, 1500.0f) * 5.0f, 1.0f);
    sky += sun_color * min(pow(sun_amount, 10.0f) * .6f, 1.0f);
    return abs(sky);
}

SBX_FN float density_func(_in(vec3) pos_in, _in(float) height, bool active) {   // :62-86, USE_NOISE_TEX (collective)
    const vec3 pos = pos_in * cld_noise_factor;
    const vec2 s = sbx_sample_both(pos, active);
    float shape = s.x;
    const float w = s.y;
    const float ww = mix(w, 1.0f - w, height);
    shape = remap(shape, ww * .7f, 1.0f, 0.0f, 1.0f);
    const float cov = 1.0f - cld_coverage;
    return shape * smoothstep(cov, cov + .0135f, shape);
}

SBX_FN float illuminate_volume(_in(vec3) origin, _in(vec3) V, _in(vec3) L, bool active) {   // :91-123 (collective)
    const float dt = cld_thick / float(cld_march_steps);
    vec3 pos = origin;
    float transmittance = 1.0f;
    pos += L * dt;                             // don't sample just where the main raymarcher is
    for (int i = 0; i < illum_march_steps; i++) {               // uniform trip count
        const float height = float(i) / float(illum_march_steps);
        const float density = density_func(pos, height, active);
        transmittance *= exp(-density * sigma_scattering * dt);
        pos += L * dt;
    }
    return transmittance * sun_power * henyey_greenstein_phase_func(clamp(dot(L, V), 0.0f, 1.0f));
}

// render_clouds (:153-202) with warp-uniform control flow: every lane of the warp walks the march loop until no lane is
// live any more; a lane that is not live (under the horizon, or past its alpha > .999 exit, :197) takes part in the
// texture-staging collectives and touches nothing else, so each pixel sees exactly the reference's sequence of operations
SBX_FN vec4 render_clouds(_in(ray_t) eye, bool live) {
    const vec3 projection = eye.direction / eye.direction.y;
    vec3 origin = eye.origin + projection * 150.0f;
    origin += wind_dir * u_time * (1.0f / cld_noise_factor);
    volume_sampler_t cloud = construct_volume(origin);
    float t = 0.0f;
    const float dt = cld_thick / float(cld_march_steps);
    for (int i = 0; i < cld_march_steps; i++) {
        if (!__any_sync(sbx_lanes, live)) break;                // uniform exit
        if (live) {
            cloud.height = float(i) / float(cld_march_steps);
            cloud.pos = cloud.origin + t * projection;
            t += dt;
        }
        const float density = density_func(cloud.pos, cloud.height, live);
        const bool in_cloud = live && !(density < .005f);       // integrate_volume, :125-148
        if (__any_sync(sbx_lanes, in_cloud)) {                  // uniform
            const float illum = illuminate_volume(cloud.pos, eye.direction, sun_dir, in_cloud);
            if (in_cloud) {
                const float T_i = exp(-density * sigma_scattering * dt);
                cloud.transmittance *= T_i;
                cloud.radiance += (density * sigma_scattering) * illum * cloud.transmittance * dt;
                cloud.alpha += (1.0f - T_i) * (1.0f - cloud.alpha);
            }
        }
        if (live && cloud.alpha > .999f) live = false;
    }
    const float cutoff = dot(eye.direction, vec3(0.0f, 1.0f, 0.0f));
    return vec4(cloud.radiance, cloud.alpha * smoothstep(.0f, .2f, cutoff));
}

SBX_FN vec3 render(_in(ray_t) eye_ray, _in(vec3) point_cam) {   // :204-218
    const vec3 sky = render_sky_color(eye_ray.direction);
    const bool below = dot(eye_ray.direction, vec3(0.0f, 1.0f, 0.0f)) < 0.05f;
    const vec4 cld = render_clouds(eye_ray, !below);             // (every lane enters: the march is warp-collective)
    if (below) return sky;
    const vec3 col = mix(sky, cld.rgb, cld.a);
    return abs(col);
}

#define FOV 1.0f   // :220
#include "main.h"
