// util.h -- camera / rotation / colour helpers (replaces src/util.h:5-138).

// look-at camera: right = up x fwd, up = fwd x right  (src/util.h:5-20)
SBX_FN ray_t get_primary_ray(_in(vec3) cam_local_point, _inout(vec3) cam_origin, _inout(vec3) cam_look_at) {
    const vec3 fwd = normalize(cam_look_at - cam_origin);
    const vec3 right = cross(vec3(0.0f, 1.0f, 0.0f), fwd);
    const vec3 up = cross(fwd, right);
    ray_t r;
    r.origin = cam_origin;
    r.direction = normalize(fwd + up * cam_local_point.y + right * cam_local_point.x);
    return r;
}

const mat3 mat3_ident = mat3(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);

SBX_FN mat3 transpose(_in(mat3) m) {   // src/util.h:25-32
    return mat3(m[0][0], m[1][0], m[2][0],
                m[0][1], m[1][1], m[2][1],
                m[0][2], m[1][2], m[2][2]);
}

// rotations take DEGREES; constructor arguments are columns (src/util.h:35-69)
SBX_FN mat2 rotate_2d(_in(float) angle_degrees) {
    const float a = radians(angle_degrees);
    const float sn = sin(a), cs = cos(a);
    return mat2(cs, -sn, sn, cs);
}
SBX_FN mat3 rotate_around_z(_in(float) angle_degrees) {
    const float a = radians(angle_degrees);
    const float sn = sin(a), cs = cos(a);
    return mat3(cs, -sn, 0.0f, sn, cs, 0.0f, 0.0f, 0.0f, 1.0f);
}
SBX_FN mat3 rotate_around_y(_in(float) angle_degrees) {
    const float a = radians(angle_degrees);
    const float sn = sin(a), cs = cos(a);
    return mat3(cs, 0.0f, sn, 0.0f, 1.0f, 0.0f, -sn, 0.0f, cs);
}
SBX_FN mat3 rotate_around_x(_in(float) angle_degrees) {
    const float a = radians(angle_degrees);
    const float sn = sin(a), cs = cos(a);
    return mat3(1.0f, 0.0f, 0.0f, 0.0f, cs, -sn, 0.0f, sn, cs);
}

// gamma 2.2 both ways (src/util.h:72-83)
SBX_FN vec3 linear_to_srgb(_in(vec3) color) {
    const float g = 1.0f / 2.2f;
    return vec3(pow(color.x, g), pow(color.y, g), pow(color.z, g));
}
SBX_FN vec3 srgb_to_linear(_in(vec3) color) {
    const float g = 2.2f;
    return vec3(pow(color.x, g), pow(color.y, g), pow(color.z, g));
}

SBX_FN vec3 faceforward(_in(vec3) N, _in(vec3) I, _in(vec3) Nref) {   // src/util.h:86-92
    return dot(Nref, I) < 0.0f ? N : -N;
}

SBX_FN float checkboard_pattern(_in(vec2) pos, _in(float) scale) {   // src/util.h:95-101
    const vec2 cell = floor(pos * scale);
    return mod(cell.x + cell.y, 2.0f);
}

SBX_FN float band(_in(float) start, _in(float) peak, _in(float) end, _in(float) t) {   // src/util.h:103-112
    return smoothstep(start, peak, t) * (1.0f - smoothstep(peak, end, t));
}

// Frisvad's basis without normalisation (src/util.h:116-125)
SBX_FN void fast_orthonormal_basis(_in(vec3) n, _out(vec3) f, _out(vec3) r) {
    const float a = 1.0f / (1.0f + n.z);
    const float b = -n.x * n.y * a;
    f = vec3(1.0f - n.x * n.x * a, b, -n.x);
    r = vec3(b, 1.0f - n.y * n.y * a, -n.y);
}

SBX_FN float remap(_in(float) original_value, _in(float) original_min, _in(float) original_max,
                   _in(float) new_min, _in(float) new_max) {   // src/util.h:127-138
    return new_min + (((original_value - original_min) / (original_max - original_min)) * (new_max - new_min));
}
