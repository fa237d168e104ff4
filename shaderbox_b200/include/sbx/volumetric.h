// volumetric.h -- phase functions and the march accumulator (replaces src/volumetric.h:5-68).
// Needs the app's `hg_g` macro (anisotropy) and PI at the point of inclusion.

// NB: the reference writes 1/(4 pi) as `1. / 4. * PI`, i.e. (1/4)*pi (src/volumetric.h:5-11); kept.
SBX_FN float isotropic_phase_func(float mu) { return 1.0f / 4.0f * PI; }

SBX_FN float rayleigh_phase_func(float mu) {   // :13-20
    return 3.0f * (1.0f + mu * mu) / (16.0f * PI);
}

// Henyey-Greenstein with the reference's (4 + pi) normalisation (src/volumetric.h:27-33)
SBX_FN float henyey_greenstein_phase_func(float mu) {
    return (1.0f - hg_g * hg_g) / ((4.0f + PI) * pow(1.0f + hg_g * hg_g - 2.0f * hg_g * mu, 1.5f));
}

#define shk_g (1.55f * hg_g - 0.55f * (hg_g * hg_g * hg_g))
SBX_FN float schlick_phase_func(float mu) {    // :35-45
    return (1.0f - shk_g * shk_g) / (4.0f * PI * (1.0f + shk_g * mu) * (1.0f + shk_g * mu));
}

struct volume_sampler_t {   // :47-54
    vec3 origin;            // ray start
    vec3 pos;               // current sample position
    float height;           // 0..1 inside the slab
    float transmittance;    // running Beer-Lambert product
    vec3 radiance;          // accumulated in-scattered light
    float alpha;
};

SBX_FN volume_sampler_t construct_volume(_in(vec3) origin) {   // :56-68
    volume_sampler_t v;
    v.origin = origin;
    v.pos = origin;
    v.height = 0.0f;
    v.transmittance = 1.0f;
    v.radiance = vec3(0.0f, 0.0f, 0.0f);
    v.alpha = 0.0f;
    return v;
}
