// light.h -- lights and the two shading models (replaces src/light.h:5-92).
#define LIGHT_POINT 1
#define LIGHT_DIR 2

struct light_t {
    int type;
    vec3 L;       // position (point) or direction (directional)
    vec3 color;
};

_mutable(light_t) lights[8];
_mutable(vec3) ambient_light = vec3(.01f, .01f, .01f);

SBX_FN vec3 get_light_direction(_in(light_t) light, _in(hit_t) P) {   // :18-27
    if (light.type == LIGHT_DIR) return light.L;
    return normalize(light.L - P.origin);
}

SBX_FN vec3 illum_blinn_phong(_in(vec3) V, _in(vec3) L, _in(hit_t) hit, _in(material_t) mat) {   // :44-62
    const vec3 diffuse = max(0.0f, dot(L, hit.normal)) * mat.base_color;
    const float spec_factor = 50.0f;
    const vec3 R = reflect(-L, hit.normal);                       // Phong lobe
    const vec3 specular = pow(max(0.0f, dot(R, V)), spec_factor) * vec3(1.0f, 1.0f, 1.0f);
    return diffuse + specular;
}

// Cook-Torrance: min-form geometry term, Beckmann distribution, Schlick Fresnel (:64-92)
SBX_FN vec3 illum_cook_torrance(_in(vec3) V, _in(vec3) L, _in(hit_t) hit, _in(material_t) mat) {
    const vec3 H = normalize(L + V);
    const float NdotL = dot(hit.normal, L);
    const float NdotH = dot(hit.normal, H);
    const float NdotV = dot(hit.normal, V);
    const float VdotH = dot(V, H);

    const float geo_a = (2.0f * NdotH * NdotV) / VdotH;
    const float geo_b = (2.0f * NdotH * NdotL) / VdotH;
    const float geo_term = min(1.0f, min(geo_a, geo_b));

    const float rough_sq = mat.roughness * mat.roughness;
    const float rough_a = 1.0f / (rough_sq * NdotH * NdotH * NdotH * NdotH);
    const float rough_exp = (NdotH * NdotH - 1.0f) / (rough_sq * NdotH * NdotH);
    const float rough_term = rough_a * exp(rough_exp);

    const float fresnel_term = fresnel_factor(1.0f, mat.ior, VdotH);

    const float specular = (geo_term * rough_term * fresnel_term) / (PI * NdotV * NdotL);
    return max(0.0f, NdotL) * (specular + mat.base_color);
}
