// light.h -- lights and the two shading models of the operator library (the names and results of src/light.h:5-92).
#define LIGHT_POINT 1
#define LIGHT_DIR 2

struct light_t {
    int type;     // LIGHT_POINT or LIGHT_DIR
    vec3 L;       // position of a point light, direction of a directional one
    vec3 color;
};

_mutable(light_t) lights[8];
_mutable(vec3) ambient_light = vec3(.01f, .01f, .01f);

// unit vector from the shaded point towards the light (:18-27)
SBX_FN vec3 get_light_direction(_in(light_t) light, _in(hit_t) P) {
    return light.type == LIGHT_DIR ? light.L : normalize(light.L - P.origin);
}

// Lambert diffuse plus a Phong lobe of exponent 50 around the mirrored light direction (:44-62)
SBX_FN vec3 illum_blinn_phong(_in(vec3) V, _in(vec3) L, _in(hit_t) hit, _in(material_t) mat) {
    const float lambert = max(0.0f, dot(L, hit.normal));
    const vec3 mirrored = reflect(-L, hit.normal);
    const float lobe = pow(max(0.0f, dot(mirrored, V)), 50.0f);
    return lambert * mat.base_color + lobe * vec3(1.0f, 1.0f, 1.0f);
}

// Cook-Torrance (:64-92) = G * D * F / (pi * N.V * N.L), term by term:
//   G  the min-form shadowing/masking term  min(1, 2 N.H N.V / V.H, 2 N.H N.L / V.H)
SBX_FN float sbx_ct_geometry(float NdotH, float NdotV, float NdotL, float VdotH) {
    const float view_side = (2.0f * NdotH * NdotV) / VdotH;
    const float light_side = (2.0f * NdotH * NdotL) / VdotH;
    return min(1.0f, min(view_side, light_side));
}
//   D  the Beckmann distribution  exp((c^2 - 1) / (m^2 c^2)) / (m^2 c^4),  c = N.H, m = roughness
SBX_FN float sbx_ct_beckmann(float roughness, float NdotH) {
    const float m2 = roughness * roughness;
    const float scale = 1.0f / (m2 * NdotH * NdotH * NdotH * NdotH);
    const float exponent = (NdotH * NdotH - 1.0f) / (m2 * NdotH * NdotH);
    return scale * exp(exponent);
}
//   F  Schlick's Fresnel between vacuum and the material's index of refraction
SBX_FN vec3 illum_cook_torrance(_in(vec3) V, _in(vec3) L, _in(hit_t) hit, _in(material_t) mat) {
    const vec3 H = normalize(L + V);                 // half vector
    const float NdotL = dot(hit.normal, L), NdotH = dot(hit.normal, H), NdotV = dot(hit.normal, V), VdotH = dot(V, H);
    const float G = sbx_ct_geometry(NdotH, NdotV, NdotL, VdotH);
    const float D = sbx_ct_beckmann(mat.roughness, NdotH);
    const float F = fresnel_factor(1.0f, mat.ior, VdotH);
    const float specular = (G * D * F) / (PI * NdotV * NdotL);
    return max(0.0f, NdotL) * (specular + mat.base_color);
}
