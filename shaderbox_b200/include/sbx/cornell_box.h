// cornell_box.h -- the Cornell-box scene tables (replaces src/cornell_box.h:9-87).
// Fills the per-pixel materials / lights tables plus six walls and three spheres.
#define num_cb_planes 6
_mutable(plane_t) cb_planes[num_cb_planes];
#define num_cb_spheres 3
_mutable(sphere_t) cb_spheres[num_cb_spheres];

SBX_FN void setup_material(_inout(material_t) mat, _in(vec3) diffuse, _in(float) metallic, _in(float) roughness) {
    mat.base_color = diffuse;
    mat.metallic = metallic;
    mat.roughness = roughness;
    mat.ior = 1.0f;
    mat.reflectivity = 0.0f;
    mat.translucency = 0.0f;
}

SBX_FN void setup_plane(_inout(plane_t) p, _in(vec3) n, _in(float) d, _in(int) mat_id) {
    p.direction = n;
    p.distance = d;
    p.material = mat_id;
}

#define cb_mat_white 1
#define cb_mat_red 2
#define cb_mat_blue 3
#define cb_mat_reflect 4
#define cb_mat_refract 5
#define cb_mat_green 6
#define cb_plane_ground 0
#define cb_plane_behind 1
#define cb_plane_front 2
#define cb_plane_ceiling 3
#define cb_plane_left 4
#define cb_plane_right 5
#define cb_plane_dist 2.f
#define cb_sphere_light 0
#define cb_sphere_left 1
#define cb_sphere_right 2

SBX_FN void setup_cornell_box() {   // :39-87
    // materials: three diffuse walls, a rough mirror and a "glass" ball (translucency unused)
    setup_material(materials[cb_mat_white], vec3(0.7913f, 0.7913f, 0.7913f), 0.0f, 0.5f);
    setup_material(materials[cb_mat_red], vec3(0.6795f, 0.0612f, 0.0529f), 0.0f, 0.5f);
    setup_material(materials[cb_mat_blue], vec3(0.1878f, 0.1274f, 0.4287f), 0.0f, 0.5f);
    setup_material(materials[cb_mat_reflect], vec3(0.95f, 0.64f, 0.54f), 1.0f, 0.1f);
    materials[cb_mat_reflect].reflectivity = 1.0f;
    setup_material(materials[cb_mat_refract], vec3(1.0f, 0.77f, 0.345f), 1.0f, 0.05f);
    materials[cb_mat_refract].reflectivity = 1.0f;
    materials[cb_mat_refract].translucency = 0.0f;
    materials[cb_mat_refract].ior = 1.333f;

    // walls: normal, offset, material
    setup_plane(cb_planes[cb_plane_ground], vec3(0.0f, -1.0f, 0.0f), 0.0f, cb_mat_white);
    setup_plane(cb_planes[cb_plane_ceiling], vec3(0.0f, 1.0f, 0.0f), 2.0f * cb_plane_dist, cb_mat_white);
    setup_plane(cb_planes[cb_plane_behind], vec3(0.0f, 0.0f, -1.0f), -cb_plane_dist, cb_mat_white);
    setup_plane(cb_planes[cb_plane_front], vec3(0.0f, 0.0f, 1.0f), cb_plane_dist, cb_mat_white);
    setup_plane(cb_planes[cb_plane_left], vec3(1.0f, 0.0f, 0.0f), cb_plane_dist, cb_mat_red);
    setup_plane(cb_planes[cb_plane_right], vec3(-1.0f, 0.0f, 0.0f), -cb_plane_dist, cb_mat_blue);

    // spheres: the emissive "lamp" above the ceiling, mirror ball, glass ball
    cb_spheres[cb_sphere_light].origin = vec3(0.0f, 2.5f * cb_plane_dist + 0.4f, 0.0f);
    cb_spheres[cb_sphere_light].radius = 1.5f;
    cb_spheres[cb_sphere_light].material = mat_debug;
    cb_spheres[cb_sphere_left].origin = vec3(0.75f, 1.0f, -0.75f);
    cb_spheres[cb_sphere_left].radius = 0.75f;
    cb_spheres[cb_sphere_left].material = cb_mat_reflect;
    cb_spheres[cb_sphere_right].origin = vec3(-0.75f, 0.75f, 0.75f);
    cb_spheres[cb_sphere_right].radius = 0.75f;
    cb_spheres[cb_sphere_right].material = cb_mat_refract;

    lights[0].type = LIGHT_POINT;
    lights[0].L = vec3(0.0f, 2.0f * cb_plane_dist - 0.2f, 0.0f);
    lights[0].color = vec3(1.0f, 1.0f, 1.0f);
}
