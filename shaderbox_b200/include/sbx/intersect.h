// intersect.h -- analytic ray/sphere and ray/plane tests of the operator library (the names and results of
// src/intersect.h:7-77).  A test records its hit only if it is nearer than what `hit` already holds.

SBX_FN void sbx_record_hit(_inout(hit_t) hit, float t, int material, _in(vec3) where, _in(vec3) normal) {
    hit.t = t;
    hit.material_id = material;
    hit.origin = where;
    hit.normal = normal;
}

// The geometric sphere solution: with m = centre - origin, the ray passes the centre at parameter
// along = m.D and at squared distance miss2 = m.m - along^2; the surface is half_chord either side of it.
struct sbx_sphere_chord {
    float along, miss2, radius2;
};
SBX_FN sbx_sphere_chord sbx_chord(_in(ray_t) ray, _in(sphere_t) sphere) {
    const vec3 m = sphere.origin - ray.origin;
    sbx_sphere_chord c;
    c.radius2 = sphere.radius * sphere.radius;
    c.along = dot(m, ray.direction);
    c.miss2 = dot(m, m) - c.along * c.along;
    return c;
}
SBX_FN void sbx_hit_sphere_at(_in(ray_t) ray, _in(sphere_t) sphere, float t, _inout(hit_t) hit) {
    const vec3 where = ray.origin + ray.direction * t;
    sbx_record_hit(hit, t, sphere.material, where, (where - sphere.origin) / sphere.radius);
}

SBX_FN void intersect_sphere(_in(ray_t) ray, _in(sphere_t) sphere, _inout(hit_t) hit) {   // :7-33
    const sbx_sphere_chord c = sbx_chord(ray, sphere);
    if (c.along < 0.0f || c.miss2 > c.radius2) return;       // centre behind the origin, or the ray passes outside
    const float half_chord = sqrt(c.radius2 - c.miss2);
    const float t_in = c.along - half_chord, t_out = c.along + half_chord;
    const float t = t_in < 0.0f ? t_out : t_in;              // origin inside the sphere: leave through the far side
    if (t > hit.t) return;
    sbx_hit_sphere_at(ray, sphere, t, hit);
}

// no rejection at all: the caller knows the origin is inside and wants the entry-side root (:35-53)
SBX_FN void intersect_sphere_from_inside(_in(ray_t) ray, _in(sphere_t) sphere, _inout(hit_t) hit) {
    const sbx_sphere_chord c = sbx_chord(ray, sphere);
    sbx_hit_sphere_at(ray, sphere, c.along - sqrt(c.radius2 - c.miss2), hit);
}

// One-sided plane (:61-77): t = (P0 - O).N / N.D with the plane point P0 = (d, d, d) as the reference has it;
// rays with N.D below 1e-6 (parallel, or arriving from the other side) never hit.
SBX_FN void intersect_plane(_in(ray_t) ray, _in(plane_t) p, _inout(hit_t) hit) {
    const float facing = dot(p.direction, ray.direction);
    if (facing < 1e-6f) return;
    const float t = dot(vec3(p.distance) - ray.origin, p.direction) / facing;
    if (t < 0.0f || t > hit.t) return;
    sbx_record_hit(hit, t, p.material, ray.origin + ray.direction * t, faceforward(p.direction, ray.direction, p.direction));
}
