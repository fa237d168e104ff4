// util_optics.h -- Fresnel / reflect / refract of the operator library (the names and results of src/util_optics.h:5-35).

// Schlick's approximation: R0 + (1 - R0)(1 - cos)^5 with R0 = ((n1 - n2) / (n1 + n2))^2.  The fifth power is the
// left-to-right product the reference writes.
SBX_FN float fresnel_factor(_in(float) n1, _in(float) n2, _in(float) VdotH) {
    const float ratio = (n1 - n2) / (n1 + n2);
    const float head_on = ratio * ratio;
    const float grazing = 1.0f - VdotH;
    float fifth = grazing * grazing;
    fifth *= grazing;
    fifth *= grazing;
    fifth *= grazing;
    return head_on + (1.0f - head_on) * fifth;
}

// mirror the incident vector about the normal (:17-22)
SBX_FN vec3 reflect(_in(vec3) incident, _in(vec3) normal) {
    const float twice = 2.0f * dot(normal, incident);
    return incident - twice * normal;
}

// Snell with relative index n; past the critical angle the ray is reflected instead (:24-35)
SBX_FN vec3 refract(_in(vec3) incident, _in(vec3) normal, _in(float) n) {
    const float cos_in = -dot(normal, incident);
    const float sin2_out = n * n * (1.0f - cos_in * cos_in);
    if (sin2_out > 1.0f) return reflect(incident, normal);
    return n * incident + (n * cos_in - sqrt(1.0f - sin2_out)) * normal;
}
