// util_optics.h -- Fresnel / reflect / refract (replaces src/util_optics.h:5-35).

// Schlick: R0 + (1-R0)(1-cos)^5
SBX_FN float fresnel_factor(_in(float) n1, _in(float) n2, _in(float) VdotH) {
    const float Rn = (n1 - n2) / (n1 + n2);
    const float R0 = Rn * Rn;
    const float F = 1.0f - VdotH;
    return R0 + (1.0f - R0) * (F * F * F * F * F);
}

SBX_FN vec3 reflect(_in(vec3) incident, _in(vec3) normal) {   // :17-22
    return incident - 2.0f * dot(normal, incident) * normal;
}

SBX_FN vec3 refract(_in(vec3) incident, _in(vec3) normal, _in(float) n) {   // :24-35
    const float cosi = -dot(normal, incident);
    const float sint2 = n * n * (1.0f - cosi * cosi);
    if (sint2 > 1.0f) return reflect(incident, normal);   // total internal reflection
    return n * incident + (n * cosi - sqrt(1.0f - sint2)) * normal;
}
