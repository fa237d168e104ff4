// sdf.h -- signed-distance combinators and primitives of the operator library (the names and results of
// src/sdf.h:5-171).  Scene functions return vec2 = (distance, material id).

// ---- combinators -------------------------------------------------------------------------------
SBX_FN vec2 op_add(_in(vec2) d1, _in(vec2) d2) { return d1.x < d2.x ? d1 : d2; }     // union keeps the nearer surface AND its material (:5-11)
SBX_FN float op_add(_in(float) d1, _in(float) d2) { return min(d1, d2); }             // :13-18
SBX_FN float op_sub(_in(float) d1, _in(float) d2) { return max(d1, -d2); }            // d1 minus the inside of d2 (:20-28)
SBX_FN float op_intersect(_in(float) d1, _in(float) d2) { return max(d1, d2); }       // :30-36

// polynomial smooth union: a and b blend over a band of width k (:38-47)
SBX_FN float op_blend(_in(float) a, _in(float) b, _in(float) k) {
    const float weight = clamp(0.5f + 0.5f * (b - a) / k, 0.0f, 1.0f);
    const float blended = mix(b, a, weight);
    return blended - k * weight * (1.0f - weight);
}

// ---- primitives --------------------------------------------------------------------------------
SBX_FN float sbx_hypot2(float a, float b) { return sqrt(a * a + b * b); }             // length of a 2-vector, a first

SBX_FN float sd_plane(_in(vec3) p, _in(vec3) n, _in(float) d) { return dot(n, p) + d; }   // :49-57
SBX_FN float sd_sphere(_in(vec3) p, _in(float) r) { return length(p) - r; }            // :59-65
SBX_FN float sd_box(_in(vec3) p, _in(vec3) b) {                                        // largest face distance, z innermost (:67-73)
    const float dx = abs(p.x) - b.x, dy = abs(p.y) - b.y, dz = abs(p.z) - b.z;
    return max(dx, max(dy, dz));
}
SBX_FN float sd_torus(_in(vec3) p, _in(float) R, _in(float) r) {                       // ring of radius R in the xy plane, tube r (:75-83)
    return sbx_hypot2(sbx_hypot2(p.x, p.y) - R, p.z) - r;
}
SBX_FN float sd_y_cylinder(_in(vec3) p, _in(float) r, _in(float) h) {                  // axis y, total height h (:85-93)
    return max(sbx_hypot2(p.x, p.z) - r, abs(p.y) - h / 2.0f);
}

// finite cylinder of radius R around the segment P0-P1: the distance to the infinite axis line, cut by the two
// end planes (which the reference places at |P1| and |P0| along the axis), then inflated by R (:95-109)
SBX_FN float sd_cylinder(_in(vec3) P, _in(vec3) P0, _in(vec3) P1, _in(float) R) {
    const vec3 axis = normalize(P1 - P0);
    const float to_axis = length(cross(axis, P - P0));
    const float beyond_far = sd_plane(P, axis, length(P1));
    const float beyond_near = sd_plane(P, -axis, -length(P0));
    return op_sub(op_sub(to_axis, beyond_far), beyond_near) - R;
}

// ---- quadratic Bezier tube (:111-159) ----------------------------------------------------------------
// The curve a-b-c is flattened into its own plane (u along c-b, v across, w the plane normal); in that plane the
// nearest point of a quadratic Bezier has the closed form of Hoppe's "Random-access vector graphics", evaluated
// on control points translated so that the query is the origin.
SBX_FN float det2(_in(vec2) a, _in(vec2) b) { return a.x * b.y - b.x * a.y; }
SBX_FN vec3 sd_bezier_get_closest(_in(vec2) b0, _in(vec2) b1, _in(vec2) b2) {
    const float a = det2(b0, b2), b = 2.0f * det2(b1, b0), d = 2.0f * det2(b2, b1);
    const float f = b * d - a * a;
    const vec2 e21 = b2 - b1, e10 = b1 - b0, e20 = b2 - b0;
    const vec2 g = 2.0f * (b * e21 + d * e10 + a * e20);          // gradient of f ...
    const vec2 g_perp = vec2(g.y, -g.x);                          // ... turned a quarter
    const vec2 foot = -f * g_perp / dot(g_perp, g_perp);
    const vec2 from_foot = b0 - foot;
    const float t = clamp((det2(from_foot, e20) + 2.0f * det2(e10, from_foot)) / (2.0f * a + b + d), 0.0f, 1.0f);
    return vec3(mix(mix(b0, b1, t), mix(b1, b2, t), t), t);       // de Casteljau at t, and t itself
}
SBX_FN vec2 sd_bezier(_in(vec3) a, _in(vec3) b, _in(vec3) c, _in(vec3) p, _in(float) thickness) {
    const vec3 w = normalize(cross(c - b, a - b));
    const vec3 u = normalize(c - b);
    const vec3 v = normalize(cross(w, u));
    const vec3 ra = a - b, rc = c - b, rp = p - b;                // everything relative to the knot b
    const vec2 a2 = vec2(dot(ra, u), dot(ra, v)), c2 = vec2(dot(rc, u), dot(rc, v));
    const vec2 q = vec2(dot(rp, u), dot(rp, v));                  // the query in the curve's plane ...
    const float off_plane = dot(rp, w);                           // ... and its height above it
    const vec3 nearest = sd_bezier_get_closest(a2 - q, vec2(0.0f, 0.0f) - q, c2 - q);
    const float in_plane2 = nearest.x * nearest.x + nearest.y * nearest.y;
    return vec2(0.85f * (sqrt(in_plane2 + off_plane * off_plane) - thickness), nearest.z);
}

// capsule of radius r around the segment a-b: distance to the clamped projection of p (:162-171)
SBX_FN float sd_capsule(_in(vec3) p, _in(vec3) a, _in(vec3) b, _in(float) r) {
    const vec3 seg = b - a;
    const float t = clamp(dot(p - a, seg) / dot(seg, seg), 0.0f, 1.0f);
    return length((seg * t + a) - p) - r;
}
