// hlsl_tex.h -- the HLSL resource types the USE_NOISE_TEX branch of src/app_clouds.h declares (:51-55) and the sampler
// behind `SampleLevel` (:69, :77), for an app header compiled with -DUSE_NOISE_TEX.  Included by def.h INSIDE the
// per-pixel app object, like the rest of the library.
//
// The reference binds these with HLSL register semantics -- `Texture3D u_tex_noise : register(t1);` -- which hlsltoy fills
// with PSSetShaderResources / PSSetSamplers (util/hlsltoy/src/hlsltoy.cpp:225-249, 497-503).  sbx_compile_app turns each
// such declaration into `T name = sbx_hlsl_bind(sbx_T, slot, (const T*)0);` in the text it hands to NVRTC (the file on disk is
// never modified): slot t1 is noise texture 0, t2 noise texture 1 of sbx_set_noise_volumes; t0 (the checkerboard
// Texture2D, only sampled under `#if 0`) and s0 (the sampler state: fixed, see below) bind to nothing.
//
// The sampler is the library's DEFINED rule (oracle/sbx_oracle.c tex_coord / tex_sample_r): D3D11 linear
// filtering, WRAP addressing, 8-bit sub-texel weights, lerp(a, b, w) = a (1 - w) + b w in fp32, x then y then z.  The
// volumes live in HBM as padded single-channel arrays (sbx_tex_params), so .g .b .a read as 0.

// (static members of the app object: the texture types below are nested classes and call them without an instance)
// one axis: padded index of the lower texel (the apron shifts indices by one) and the 8-bit weight
static SBX_FN void sbx_tex_axis(float u, float n, int& i0, float& w) {
    const float uw = u - floor(u);            // WRAP
    const float t = uw * n - 0.5f;            // texel space, texel centres at i + 0.5
    const float fl = floor(t);
    const float f = t - fl;
    i0 = int(fl) + 1;
    w = floor(f * 256.0f + 0.5f) / 256.0f;    // D3D11_SUBTEXEL_FRACTIONAL_BIT_COUNT = 8
}
static SBX_FN float sbx_tex_lerp(float a, float b, float w) { return a * (1.0f - w) + b * w; }
static SBX_FN float sbx_tex_blend(const float* c, float wx, float wy, float wz) {   // c: x0y0z0 x1y0z0 x0y1z0 x1y1z0 x0y0z1 ...
    const float c00 = sbx_tex_lerp(c[0], c[1], wx), c10 = sbx_tex_lerp(c[2], c[3], wx);
    const float c01 = sbx_tex_lerp(c[4], c[5], wx), c11 = sbx_tex_lerp(c[6], c[7], wx);
    return sbx_tex_lerp(sbx_tex_lerp(c00, c10, wy), sbx_tex_lerp(c01, c11, wy), wz);
}

struct SamplerState {};                       // MIN_MAG_MIP_LINEAR + WRAP (hlsltoy.cpp:244-249) is the only sampler there is

struct Texture3D {
    const sbx_tex_params* T;
    int index;                                // which of the two noise volumes; < 0: unbound
    // one lane, eight read-only loads of the padded volume (the hand-written kernel stages texel boxes with TMA instead)
    SBX_FN vec4 SampleLevel(const SamplerState&, const vec3& pos, float) const {
        if (index < 0 || T == nullptr) return vec4(0.0f, 0.0f, 0.0f, 0.0f);
        const float n = float(T->size);
        int ix, iy, iz;
        float wx, wy, wz;
        sbx_tex_axis(pos.x, n, ix, wx);
        sbx_tex_axis(pos.y, n, iy, wy);
        sbx_tex_axis(pos.z, n, iz, wz);
        const float* vol = T->vol[index];
        const size_t at = (size_t)iz * T->pitch_xy + (size_t)iy * T->pitch_x + ix;
        float c[8];
        _Pragma("unroll") for (int k = 0; k < 8; ++k)
            c[k] = __ldg(vol + at + (k & 1) + (size_t)((k >> 1) & 1) * T->pitch_x + (size_t)(k >> 2) * T->pitch_xy);
        return vec4(sbx_tex_blend(c, wx, wy, wz), 0.0f, 0.0f, 0.0f);
    }
};
struct Texture2D {                            // declared by the app (:52), sampled only under `#if 0` (:169-172)
    SBX_FN vec4 Sample(const SamplerState&, const vec2&) const { return vec4(0.0f, 0.0f, 0.0f, 0.0f); }
};

// `: register(xN)` of an HLSL declaration, as rewritten by sbx_compile_app: overloads on a null pointer of the declared type
static SBX_FN Texture3D sbx_hlsl_bind(const sbx_tex_params* T, int slot, const Texture3D*) {
    Texture3D t;
    t.T = T;
    t.index = (slot == 1 || slot == 2) ? slot - 1 : -1;
    return t;
}
static SBX_FN Texture2D sbx_hlsl_bind(const sbx_tex_params*, int, const Texture2D*) { return Texture2D(); }
static SBX_FN SamplerState sbx_hlsl_bind(const sbx_tex_params*, int, const SamplerState*) { return SamplerState(); }
