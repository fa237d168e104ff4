// main.h -- the per-pixel entry (replaces src/main.h:6-53).  Needs from the app: setup_camera,
// setup_scene, render and the FOV macro.  The pixel LOOP that calls this lives in
// sbx_kernel.cuh (one thread per pixel); this is only the body.
SBX_FN void mainImage(vec4& fragColor, const vec2& fragCoord) {
    // raster [0..res] -> NDC [0..1] -> camera plane [-aspect*fov, +aspect*fov] x [-fov, fov], z = -1
    const vec2 aspect_ratio = vec2(u_res.x / u_res.y, 1.0f);

    vec3 eye, look_at;
    setup_camera(eye, look_at);
    setup_scene();

    const vec2 point_ndc = fragCoord.xy / u_res.xy;
    const vec3 point_cam = vec3((2.0f * point_ndc - 1.0f) * aspect_ratio * FOV, -1.0f);

    const ray_t ray = get_primary_ray(point_cam, eye, look_at);
    const vec3 color = render(ray, point_cam);

#ifdef SBX_APP_ENCODE      // a hand-written scene kernel may spread the three pows over the lanes that share the pixel
    fragColor = vec4(SBX_APP_ENCODE(color), 1.0f);
#else
    fragColor = vec4(linear_to_srgb(color), 1.0f);
#endif
}
