// noise_worley.h -- tiled 3-D cellular noise (replaces src/noise_worley.h:5-51).
// Returns (F1, F2, cell id); `domain_repeat` is both the frequency and the tiling period.

SBX_FN vec3 hash_w(_in(vec3) x) {   // three decorrelated sine hashes (:5-17)
    const vec3 q = vec3(dot(x, vec3(127.1f, 311.7f, 74.7f)),
                        dot(x, vec3(269.5f, 183.3f, 246.1f)),
                        dot(x, vec3(113.5f, 271.9f, 124.6f)));
    return fract(sin(q) * 43758.5453123f);
}

SBX_FN vec3 noise_w(_in(vec3) pos, _in(float) domain_repeat) {   // :20-51
    const vec3 x = pos * domain_repeat;
    const vec3 p = floor(x);
    const vec3 f = fract(x);

    float id = 0.0f;
    vec2 res = vec2(100.0f, 100.0f);      // squared distances: nearest, second nearest
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                const vec3 b = vec3(float(i), float(j), float(k));
                const vec3 r = b - f + hash_w(mod(p + b, domain_repeat));
                const float d = dot(r, r);
                if (d < res.x) {
                    id = dot(p + b, vec3(1.0f, 57.0f, 113.0f));
                    res = vec2(d, res.x);
                } else if (d < res.y) {
                    res.y = d;
                }
            }
    return vec3(sqrt(res), abs(id));
}
