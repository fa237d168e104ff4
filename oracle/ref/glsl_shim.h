// glsl_shim.h -- TEST INFRASTRUCTURE (oracle).  Never linked into the product library.
//
// The reference's C++ build gets vec2/vec3/vec4/mat2/mat3 and the GLSL builtins from the author's
// VML library (README.md:9-10), which is NOT in /root/reference (it is neither a submodule,
// .gitmodules:1-6, nor version-pinned).  This header restates only that layer (L-1 in SURVEY.md §1)
// so that the VERBATIM reference headers under /root/reference/src compile with g++.
// Arithmetic definitions follow the GLSL 4.x spec formulas (the headers are GLSL first,
// README.md:7); where the spec gives no formula the choice is written next to the function and
// is the SAME choice the device library (shaderbox_b200/include/sbx/sbx_vec.cuh) and the C
// restatement (oracle/sbx_oracle.c) make.  Transcendentals are the platform libm (glibc 2.39),
// exactly what a C++ build of the reference calls.
//
// Compile with: g++ -O2 -std=c++17 -fsingle-precision-constant -ffp-contract=off  (never -Ofast)
#pragma once
#include <cmath>
#ifndef GLSL_COUNT
#define GLSL_COUNT(what) /* optional per-call counter hook, defined by the includer */
#endif

namespace glsl {

struct vec2; struct vec3; struct vec4;

// ---- swizzle proxies: read-only views living in a union with the component array ----------
template <int N, int A, int B> struct swz2 {
    float v[N];
    operator vec2() const;
};
template <int N, int A, int B, int C> struct swz3 {
    float v[N];
    operator vec3() const;
};

struct vec2 {
    union {
        float v[2];
        struct { float x, y; };
        struct { float r, g; };
        swz2<2, 0, 1> xy; swz2<2, 1, 0> yx; swz2<2, 0, 0> xx; swz2<2, 1, 1> yy;
    };
    vec2() : x(0), y(0) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(float a) : x(a), y(a) {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

struct vec3 {
    union {
        float v[3];
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<3, 0, 1> xy; swz2<3, 0, 2> xz; swz2<3, 1, 2> yz; swz2<3, 1, 0> yx; swz2<3, 2, 0> zx;
        swz2<3, 2, 1> zy;
        swz3<3, 0, 1, 2> xyz; swz3<3, 0, 2, 2> xzz; swz3<3, 2, 0, 2> zxz; swz3<3, 2, 2, 0> zzx;
        swz3<3, 0, 2, 1> xzy; swz3<3, 1, 0, 2> yxz; swz3<3, 1, 2, 0> yzx; swz3<3, 2, 0, 1> zxy;
        swz3<3, 2, 1, 0> zyx; swz3<3, 0, 1, 2> rgb;
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
    vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    vec3(float a, const vec2& b) : x(a), y(b.x), z(b.y) {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

struct vec4 {
    union {
        float v[4];
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<4, 0, 1> xy; swz2<4, 2, 3> zw; swz2<4, 0, 2> xz;
        swz3<4, 0, 1, 2> xyz; swz3<4, 0, 1, 2> rgb;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    vec4(const vec3& a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
    vec4(const vec2& a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};

template <int N, int A, int B> inline swz2<N, A, B>::operator vec2() const { return vec2(v[A], v[B]); }
template <int N, int A, int B, int C> inline swz3<N, A, B, C>::operator vec3() const {
    return vec3(v[A], v[B], v[C]);
}

// ---- componentwise arithmetic -----------------------------------------------------------
#define GLSL_VEC_OPS(V, ...)                                                                      \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] + b.v[i]; return r; } \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] - b.v[i]; return r; } \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] * b.v[i]; return r; } \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] / b.v[i]; return r; } \
    inline V operator+(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] + b; return r; }        \
    inline V operator-(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] - b; return r; }        \
    inline V operator*(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] * b; return r; }        \
    inline V operator/(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a.v[i] / b; return r; }        \
    inline V operator+(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a + b.v[i]; return r; }        \
    inline V operator-(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a - b.v[i]; return r; }        \
    inline V operator*(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a * b.v[i]; return r; }        \
    inline V operator/(float a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = a / b.v[i]; return r; }        \
    inline V operator-(const V& a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = -a.v[i]; return r; }                    \
    inline V& operator+=(V& a, const V& b) { a = a + b; return a; }                               \
    inline V& operator-=(V& a, const V& b) { a = a - b; return a; }                               \
    inline V& operator*=(V& a, const V& b) { a = a * b; return a; }                               \
    inline V& operator/=(V& a, const V& b) { a = a / b; return a; }                               \
    inline V& operator+=(V& a, float b) { a = a + b; return a; }                                  \
    inline V& operator-=(V& a, float b) { a = a - b; return a; }                                  \
    inline V& operator*=(V& a, float b) { a = a * b; return a; }                                  \
    inline V& operator/=(V& a, float b) { a = a / b; return a; }
GLSL_VEC_OPS(vec2, 2)
GLSL_VEC_OPS(vec3, 3)
GLSL_VEC_OPS(vec4, 4)
#undef GLSL_VEC_OPS

// ---- scalar builtins --------------------------------------------------------------------
// min/max are IEEE minNum/maxNum (what GPUs implement; SURVEY.md §8 hazards: app_planet.h:270-273).
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float abs(float a) { return ::fabsf(a); }
inline float floor(float a) { return ::floorf(a); }
inline float fract(float a) { return a - ::floorf(a); }                       // GLSL: x - floor(x)
inline float mod(float a, float b) { return a - b * ::floorf(a / b); }        // GLSL: x - y*floor(x/y)
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }  // GLSL: min(max(x,lo),hi)
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }   // GLSL: x(1-a)+ya
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) {                        // GLSL spec formula
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float radians(float d) { return d * 0.017453292519943295f; }           // d * fl(pi/180)
inline float sqrt(float a) { GLSL_COUNT(sqrt_); return ::sqrtf(a); }
inline float sin(float a) { GLSL_COUNT(sin_); return ::sinf(a); }
inline float cos(float a) { GLSL_COUNT(cos_); return ::cosf(a); }
inline float tan(float a) { GLSL_COUNT(other_); return ::tanf(a); }
inline float exp(float a) { GLSL_COUNT(exp_); return ::expf(a); }
inline float pow(float a, float b) { GLSL_COUNT(pow_); return ::powf(a, b); }
inline float acos(float a) { GLSL_COUNT(other_); return ::acosf(a); }
inline float atan(float y, float x) { GLSL_COUNT(other_); return ::atan2f(y, x); }

// ---- vector builtins ---------------------------------------------------------------------
#define GLSL_MAP1(F, V, N) inline V F(const V& a) { V r; for (int i = 0; i < N; ++i) r.v[i] = F(a.v[i]); return r; }
#define GLSL_MAP1_ALL(F) GLSL_MAP1(F, vec2, 2) GLSL_MAP1(F, vec3, 3) GLSL_MAP1(F, vec4, 4)
GLSL_MAP1_ALL(abs) GLSL_MAP1_ALL(floor) GLSL_MAP1_ALL(fract) GLSL_MAP1_ALL(sqrt)
GLSL_MAP1_ALL(sin) GLSL_MAP1_ALL(cos) GLSL_MAP1_ALL(exp)
#undef GLSL_MAP1_ALL
#undef GLSL_MAP1
#define GLSL_VEC_FUNCS(V, N)                                                                      \
    inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.v[i] = min(a.v[i], b.v[i]); return r; } \
    inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.v[i] = max(a.v[i], b.v[i]); return r; } \
    inline V min(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r.v[i] = min(a.v[i], b); return r; }         \
    inline V max(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r.v[i] = max(a.v[i], b); return r; }         \
    inline V mod(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r.v[i] = mod(a.v[i], b); return r; }         \
    inline V mod(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.v[i] = mod(a.v[i], b.v[i]); return r; } \
    inline V pow(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.v[i] = pow(a.v[i], b.v[i]); return r; } \
    inline V clamp(const V& a, float lo, float hi) { V r; for (int i = 0; i < N; ++i) r.v[i] = clamp(a.v[i], lo, hi); return r; } \
    inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < N; ++i) r.v[i] = mix(a.v[i], b.v[i], t); return r; } \
    inline V mix(const V& a, const V& b, const V& t) { V r; for (int i = 0; i < N; ++i) r.v[i] = mix(a.v[i], b.v[i], t.v[i]); return r; } \
    inline V step(float e, const V& a) { V r; for (int i = 0; i < N; ++i) r.v[i] = step(e, a.v[i]); return r; }       \
    inline V smoothstep(float e0, float e1, const V& a) { V r; for (int i = 0; i < N; ++i) r.v[i] = smoothstep(e0, e1, a.v[i]); return r; }
GLSL_VEC_FUNCS(vec2, 2)
GLSL_VEC_FUNCS(vec3, 3)
GLSL_VEC_FUNCS(vec4, 4)
#undef GLSL_VEC_FUNCS

// dot: products summed left to right, ((x+y)+z)+w
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
// length = sqrt(dot(v,v)); normalize = v / length(v) (three IEEE divisions)
inline float length(const vec2& a) { return sqrt(dot(a, a)); }
inline float length(const vec3& a) { return sqrt(dot(a, a)); }
inline float length(const vec4& a) { return sqrt(dot(a, a)); }
inline vec2 normalize(const vec2& a) { return a / length(a); }
inline vec3 normalize(const vec3& a) { return a / length(a); }
inline vec4 normalize(const vec4& a) { return a / length(a); }
inline float distance(const vec3& a, const vec3& b) { return length(a - b); }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// ---- matrices: column-major, m[col][row], constructor takes columns ---------------------------
struct mat2 {
    vec2 c[2];
    mat2() {}
    mat2(float a, float b, float d, float e) { c[0] = vec2(a, b); c[1] = vec2(d, e); }
    vec2& operator[](int i) { return c[i]; }
    const vec2& operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    mat3() {}
    mat3(float a, float b, float cc, float d, float e, float f, float g, float h, float i) {
        c[0] = vec3(a, b, cc); c[1] = vec3(d, e, f); c[2] = vec3(g, h, i);
    }
    mat3(const vec3& a, const vec3& b, const vec3& cc) { c[0] = a; c[1] = b; c[2] = cc; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
// M*v = c0*v.x + c1*v.y + c2*v.z  (accumulated left to right)
inline vec2 operator*(const mat2& m, const vec2& v) { return m.c[0] * v.x + m.c[1] * v.y; }
inline vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
// v*M = (dot(v,c0), dot(v,c1), dot(v,c2))
inline vec2 operator*(const vec2& v, const mat2& m) { return vec2(dot(v, m.c[0]), dot(v, m.c[1])); }
inline vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
inline mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
inline mat2 operator*(const mat2& a, const mat2& b) {
    mat2 r; r.c[0] = a * b.c[0]; r.c[1] = a * b.c[1]; return r;
}

// GLSL builtins the vector library hands to a C++ build.  src/util_optics.h:16-36 supplies reflect/refract
// itself (as members, which shadow these) for the apps that include it; src/app_vinyl.h:316 calls reflect
// WITHOUT including util_optics.h, i.e. it relies on the host library's.  GLSL 4.x spec: I - 2*dot(N,I)*N,
// the same expression as src/util_optics.h:20.
inline vec3 reflect(const vec3& incident, const vec3& normal) { return incident - 2.0f * dot(normal, incident) * normal; }

}  // namespace glsl
