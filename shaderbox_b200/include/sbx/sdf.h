// sdf.h -- signed-distance operators and primitives (replaces src/sdf.h:5-171).
// vec2 results carry (distance, material id).

SBX_FN vec2 op_add(_in(vec2) d1, _in(vec2) d2) { return d1.x < d2.x ? d1 : d2; }          // union :5-11
SBX_FN float op_add(_in(float) d1, _in(float) d2) { return min(d1, d2); }                  // :13-18
SBX_FN float op_sub(_in(float) d1, _in(float) d2) { return max(d1, -d2); }                 // carve d2 out of d1 :20-28
SBX_FN float op_intersect(_in(float) d1, _in(float) d2) { return max(d1, d2); }            // :30-36

// polynomial smooth minimum, k = blend radius (src/sdf.h:38-47)
SBX_FN float op_blend(_in(float) a, _in(float) b, _in(float) k) {
    const float h = clamp(0.5f + 0.5f * (b - a) / k, 0.0f, 1.0f);
    return mix(b, a, h) - k * h * (1.0f - h);
}

SBX_FN float sd_plane(_in(vec3) p, _in(vec3) n, _in(float) d) { return dot(n, p) + d; }     // :49-57
SBX_FN float sd_sphere(_in(vec3) p, _in(float) r) { return length(p) - r; }                // :59-65
SBX_FN float sd_box(_in(vec3) p, _in(vec3) b) {                                            // :67-73
    return max(abs(p.x) - b.x, max(abs(p.y) - b.y, abs(p.z) - b.z));
}
SBX_FN float sd_torus(_in(vec3) p, _in(float) R, _in(float) r) {                           // ring in xy :75-83
    const float ring = sqrt(p.x * p.x + p.y * p.y) - R;
    return sqrt(ring * ring + p.z * p.z) - r;
}
SBX_FN float sd_y_cylinder(_in(vec3) p, _in(float) r, _in(float) h) {                      // :85-93
    return max(sqrt(p.x * p.x + p.z * p.z) - r, abs(p.y) - h / 2.0f);
}

// capped cylinder along P0->P1 of thickness R (src/sdf.h:95-109)
SBX_FN float sd_cylinder(_in(vec3) P, _in(vec3) P0, _in(vec3) P1, _in(float) R) {
    const vec3 axis = normalize(P1 - P0);
    const float radial = length(cross(axis, P - P0));
    const float cap_far = sd_plane(P, axis, length(P1));
    const float cap_near = sd_plane(P, -axis, -length(P0));
    return op_sub(op_sub(radial, cap_far), cap_near) - R;
}

// quadratic Bezier tube (Hoppe's closest point in the curve's plane; src/sdf.h:114-159)
SBX_FN float det2(_in(vec2) a, _in(vec2) b) { return a.x * b.y - b.x * a.y; }
SBX_FN vec3 sd_bezier_get_closest(_in(vec2) b0, _in(vec2) b1, _in(vec2) b2) {
    const float a = det2(b0, b2);
    const float b = 2.0f * det2(b1, b0);
    const float d = 2.0f * det2(b2, b1);
    const float f = b * d - a * a;
    const vec2 d21 = b2 - b1;
    const vec2 d10 = b1 - b0;
    const vec2 d20 = b2 - b0;
    vec2 gf = 2.0f * (b * d21 + d * d10 + a * d20);
    gf = vec2(gf.y, -gf.x);
    const vec2 pp = -f * gf / dot(gf, gf);
    const vec2 d0p = b0 - pp;
    const float ap = det2(d0p, d20);
    const float bp = 2.0f * det2(d10, d0p);
    const float t = clamp((ap + bp) / (2.0f * a + b + d), 0.0f, 1.0f);
    return vec3(mix(mix(b0, b1, t), mix(b1, b2, t), t), t);
}
SBX_FN vec2 sd_bezier(_in(vec3) a, _in(vec3) b, _in(vec3) c, _in(vec3) p, _in(float) thickness) {
    const vec3 w = normalize(cross(c - b, a - b));
    const vec3 u = normalize(c - b);
    const vec3 v = normalize(cross(w, u));

    const vec2 a2 = vec2(dot(a - b, u), dot(a - b, v));
    const vec2 b2 = vec2(0.0f, 0.0f);
    const vec2 c2 = vec2(dot(c - b, u), dot(c - b, v));
    const vec3 p3 = vec3(dot(p - b, u), dot(p - b, v), dot(p - b, w));
    const vec2 q = vec2(p3.x, p3.y);

    const vec3 cp = sd_bezier_get_closest(a2 - q, b2 - q, c2 - q);
    return vec2(0.85f * (sqrt(cp.x * cp.x + cp.y * cp.y + p3.z * p3.z) - thickness), cp.z);
}

SBX_FN float sd_capsule(_in(vec3) p, _in(vec3) a, _in(vec3) b, _in(float) r) {              // :162-171
    const vec3 ab = b - a;
    const float t = clamp(dot(p - a, ab) / dot(ab, ab), 0.0f, 1.0f);
    return length((ab * t + a) - p) - r;
}
