// fbm.h -- fractional-Brownian-motion generators (replaces src/fbm.h:6,8).
// `_basis` is an expression over the local variable `p` (and `L` for the tiled form).
#define DECL_FBM_FUNC(_name, _octaves, _basis)                                                   \
    SBX_FN float _name(_in(vec3) pos, _in(float) lacunarity, _in(float) init_gain, _in(float) gain) { \
        vec3 p = pos;                                                                            \
        float H = init_gain;                                                                     \
        float t = 0.0f;                                                                          \
        _Pragma("unroll") for (int i = 0; i < _octaves; i++) {                                   \
            t += _basis * H;                                                                     \
            p *= lacunarity;                                                                     \
            H *= gain;                                                                           \
        }                                                                                        \
        return t;                                                                                \
    }

#define DECL_FBM_FUNC_TILE(_name, _octaves, _basis)                                              \
    SBX_FN float _name(_in(vec3) pos, _in(float) lacunarity, _in(float) init_gain, _in(float) gain) { \
        vec3 p = pos;                                                                            \
        float H = init_gain;                                                                     \
        float L = lacunarity;                                                                    \
        float t = 0.0f;                                                                          \
        _Pragma("unroll") for (int i = 0; i < _octaves; i++) {                                   \
            t += _basis * H;                                                                     \
            L *= lacunarity;                                                                     \
            H *= gain;                                                                           \
        }                                                                                        \
        return t;                                                                                \
    }
