// intersect.h -- analytic ray/sphere and ray/plane tests (replaces src/intersect.h:7-77).
// All three update `hit` only when they find something nearer than hit.t.

SBX_FN void intersect_sphere(_in(ray_t) ray, _in(sphere_t) sphere, _inout(hit_t) hit) {   // :7-33
    const vec3 rc = sphere.origin - ray.origin;
    const float radius2 = sphere.radius * sphere.radius;
    const float tca = dot(rc, ray.direction);
    if (tca < 0.0f) return;                       // centre behind the ray

    const float d2 = dot(rc, rc) - tca * tca;
    if (d2 > radius2) return;                     // passes outside

    const float thc = sqrt(radius2 - d2);
    float t0 = tca - thc;
    const float t1 = tca + thc;
    if (t0 < 0.0f) t0 = t1;                       // origin inside: take the far root
    if (t0 > hit.t) return;

    const vec3 impact = ray.origin + ray.direction * t0;
    hit.t = t0;
    hit.material_id = sphere.material;
    hit.origin = impact;
    hit.normal = (impact - sphere.origin) / sphere.radius;
}

SBX_FN void intersect_sphere_from_inside(_in(ray_t) ray, _in(sphere_t) sphere, _inout(hit_t) hit) {   // :35-53
    const vec3 rc = sphere.origin - ray.origin;
    const float radius2 = sphere.radius * sphere.radius;
    const float tca = dot(rc, ray.direction);
    const float d2 = dot(rc, rc) - tca * tca;
    const float thc = sqrt(radius2 - d2);
    const float t0 = tca - thc;
    const vec3 impact = ray.origin + ray.direction * t0;
    hit.t = t0;
    hit.material_id = sphere.material;
    hit.origin = impact;
    hit.normal = (impact - sphere.origin) / sphere.radius;
}

// one-sided: rays with N.D < 1e-6 are rejected; the plane point is (d,d,d) as in the reference (:61-77)
SBX_FN void intersect_plane(_in(ray_t) ray, _in(plane_t) p, _inout(hit_t) hit) {
    const float denom = dot(p.direction, ray.direction);
    if (denom < 1e-6f) return;

    const vec3 P0 = vec3(p.distance, p.distance, p.distance);
    const float t = dot(P0 - ray.origin, p.direction) / denom;
    if (t < 0.0f || t > hit.t) return;

    hit.t = t;
    hit.material_id = p.material;
    hit.origin = ray.origin + ray.direction * t;
    hit.normal = faceforward(p.direction, ray.direction, p.direction);
}
