// fbm.h -- fractional-Brownian-motion generators (replaces src/fbm.h:6,8).
// `_basis` is an expression over the local variable `p` (and `L` for the tiled form).
// Octave loops of up to 4 octaves are unrolled; longer ones stay loops: APP_PLANET's 7-octave normal
// fbms, inlined 12 times, otherwise push the kernel past the instruction cache (195 KB of SASS,
// "no instruction" the top stall -- profiles/r01f).
#define SBX_PRAGMA(x) _Pragma(#x)
#define DECL_FBM_FUNC(_name, _octaves, _basis)                                                   \
    SBX_FN float _name(_in(vec3) pos, _in(float) lacunarity, _in(float) init_gain, _in(float) gain) { \
        vec3 p = pos;                                                                            \
        float H = init_gain;                                                                     \
        float t = 0.0f;                                                                          \
        SBX_PRAGMA(unroll (_octaves <= 4 ? _octaves : 1)) for (int i = 0; i < _octaves; i++) {                                   \
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
        SBX_PRAGMA(unroll (_octaves <= 4 ? _octaves : 1)) for (int i = 0; i < _octaves; i++) {                                   \
            t += _basis * H;                                                                     \
            L *= lacunarity;                                                                     \
            H *= gain;                                                                           \
        }                                                                                        \
        return t;                                                                                \
    }
