// noise_worley.h -- tiled 3-D cellular noise of the operator library (the names and results of
// src/noise_worley.h:5-51).  noise_w returns (F1, F2, cell id): the distances to the nearest and second nearest
// feature point and the id of the nearest one's cell; `domain_repeat` is both the frequency and the tiling period.

// feature-point offset of a lattice cell: three decorrelated sine hashes (:5-17)
SBX_FN vec3 hash_w(_in(vec3) x) {
    const vec3 phase = vec3(dot(x, vec3(127.1f, 311.7f, 74.7f)),
                            dot(x, vec3(269.5f, 183.3f, 246.1f)),
                            dot(x, vec3(113.5f, 271.9f, 124.6f)));
    return fract(sin(phase) * 43758.5453123f);
}

// the running two-nearest search over the 27 neighbouring cells (:20-51)
struct sbx_worley_state {
    float nearest2, second2, id;      // squared distances and the nearest cell's id
};
SBX_FN void sbx_worley_visit(_inout(sbx_worley_state) w, _in(vec3) cell, _in(vec3) offset, _in(vec3) frac, float domain_repeat) {
    const vec3 to_feature = offset - frac + hash_w(mod(cell + offset, domain_repeat));   // wrapped lattice -> tileable
    const float d2 = dot(to_feature, to_feature);
    if (d2 < w.nearest2) {
        w.id = dot(cell + offset, vec3(1.0f, 57.0f, 113.0f));
        w.second2 = w.nearest2;
        w.nearest2 = d2;
    } else if (d2 < w.second2) {
        w.second2 = d2;
    }
}
SBX_FN vec3 noise_w(_in(vec3) pos, _in(float) domain_repeat) {
    const vec3 x = pos * domain_repeat;
    const vec3 cell = floor(x), frac = fract(x);
    sbx_worley_state w;
    w.nearest2 = 100.0f; w.second2 = 100.0f; w.id = 0.0f;
    for (int k = -1; k <= 1; k++)             // z outermost, x innermost: ties keep the first cell met
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) sbx_worley_visit(w, cell, vec3(float(i), float(j), float(k)), frac, domain_repeat);
    return vec3(sqrt(w.nearest2), sqrt(w.second2), abs(w.id));
}
