// util.h -- camera, rotations, colour and small scalar helpers of the operator library
// (the names and results of src/util.h:5-138; every function notes the lines it stands in for).

// Camera (src/util.h:5-20).  The look-at frame is (right, up, fwd) with right = Y x fwd and up = fwd x right;
// the primary ray leaves the eye through the camera-plane point (x along right, y along up, fwd at unit depth).
struct sbx_camera_frame {
    vec3 right, up, fwd;
};
SBX_FN sbx_camera_frame sbx_look_at(_in(vec3) eye, _in(vec3) target) {
    sbx_camera_frame f;
    f.fwd = normalize(target - eye);
    f.right = cross(vec3(0.0f, 1.0f, 0.0f), f.fwd);
    f.up = cross(f.fwd, f.right);
    return f;
}
SBX_FN ray_t get_primary_ray(_in(vec3) cam_local_point, _inout(vec3) cam_origin, _inout(vec3) cam_look_at) {
    const sbx_camera_frame f = sbx_look_at(cam_origin, cam_look_at);
    ray_t ray;
    ray.origin = cam_origin;
    ray.direction = normalize(f.fwd + f.up * cam_local_point.y + f.right * cam_local_point.x);   // fwd, then up, then right
    return ray;
}

const mat3 mat3_ident = mat3(vec3(1.0f, 0.0f, 0.0f), vec3(0.0f, 1.0f, 0.0f), vec3(0.0f, 0.0f, 1.0f));

SBX_FN mat3 transpose(_in(mat3) m) {   // rows become columns (src/util.h:25-32)
    mat3 t;
    _Pragma("unroll") for (int col = 0; col < 3; ++col) t.c[col] = vec3(m.c[0].v[col], m.c[1].v[col], m.c[2].v[col]);
    return t;
}

// Rotations (src/util.h:35-69): angles in DEGREES, matrices given column by column.  Every builder needs the
// sine AND the cosine of its angle, so both come from one argument reduction (sbx_sincos).
SBX_FN sbx_sincos_t sbx_rotation(float angle_degrees) { return sbx_sincos(radians(angle_degrees)); }
SBX_FN mat2 rotate_2d(_in(float) angle_degrees) {
    const sbx_sincos_t r = sbx_rotation(angle_degrees);
    return mat2(r.c, -r.s, r.s, r.c);
}
SBX_FN mat3 rotate_around_z(_in(float) angle_degrees) {
    const sbx_sincos_t r = sbx_rotation(angle_degrees);
    return mat3(vec3(r.c, -r.s, 0.0f), vec3(r.s, r.c, 0.0f), vec3(0.0f, 0.0f, 1.0f));
}
SBX_FN mat3 rotate_around_y(_in(float) angle_degrees) {
    const sbx_sincos_t r = sbx_rotation(angle_degrees);
    return mat3(vec3(r.c, 0.0f, r.s), vec3(0.0f, 1.0f, 0.0f), vec3(-r.s, 0.0f, r.c));
}
SBX_FN mat3 rotate_around_x(_in(float) angle_degrees) {
    const sbx_sincos_t r = sbx_rotation(angle_degrees);
    return mat3(vec3(1.0f, 0.0f, 0.0f), vec3(0.0f, r.c, -r.s), vec3(0.0f, r.s, r.c));
}

// Display gamma 2.2, applied per channel (src/util.h:72-83)
SBX_FN vec3 sbx_gamma(_in(vec3) c, float exponent) { return vec3(pow(c.x, exponent), pow(c.y, exponent), pow(c.z, exponent)); }
SBX_FN vec3 linear_to_srgb(_in(vec3) color) { return sbx_gamma(color, 1.0f / 2.2f); }
SBX_FN vec3 srgb_to_linear(_in(vec3) color) { return sbx_gamma(color, 2.2f); }

// N if the reference normal faces against I, else -N (src/util.h:86-92)
SBX_FN vec3 faceforward(_in(vec3) N, _in(vec3) I, _in(vec3) Nref) {
    const bool facing = dot(Nref, I) < 0.0f;
    return facing ? N : -N;
}

// 0/1 checkerboard of cell size 1/scale (src/util.h:95-101)
SBX_FN float checkboard_pattern(_in(vec2) pos, _in(float) scale) {
    const vec2 cell = floor(pos * scale);
    return mod(cell.x + cell.y, 2.0f);
}

// rises over [start, peak], falls over [peak, end]: the product of two smoothsteps (src/util.h:103-112)
SBX_FN float band(_in(float) start, _in(float) peak, _in(float) end, _in(float) t) {
    const float rise = smoothstep(start, peak, t);
    const float fall = smoothstep(peak, end, t);
    return rise * (1.0f - fall);
}

// Frisvad's branch-free tangent frame, not normalised (src/util.h:116-125)
SBX_FN void fast_orthonormal_basis(_in(vec3) n, _out(vec3) f, _out(vec3) r) {
    const float inv = 1.0f / (1.0f + n.z);
    const float shear = -n.x * n.y * inv;
    f = vec3(1.0f - n.x * n.x * inv, shear, -n.x);
    r = vec3(shear, 1.0f - n.y * n.y * inv, -n.y);
}

// affine map of [original_min, original_max] onto [new_min, new_max] (src/util.h:127-138)
SBX_FN float remap(_in(float) original_value, _in(float) original_min, _in(float) original_max,
                   _in(float) new_min, _in(float) new_max) {
    const float unit = (original_value - original_min) / (original_max - original_min);
    return new_min + unit * (new_max - new_min);
}
