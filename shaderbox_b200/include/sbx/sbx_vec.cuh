// sbx_vec.cuh -- the vector layer of the device operator library (what VML gives the reference's
// C++ build, README.md:9-10; what GLSL gives its GPU builds).  vec2/3/4, mat2/3, swizzles and the
// GLSL builtins the shaderbox headers are written against (census: SURVEY.md §8c).
//
// Arithmetic contract (identical in oracle/ref/glsl_shim.h and oracle/sbx_oracle.c):
//   every + - * / sqrt is one IEEE-754 binary32 operation, never contracted (compile with
//   --fmad=false), in the order written here; dot sums left to right; length = sqrt(dot);
//   normalize = v / length(v); mix = x*(1-a) + y*a; clamp = min(max(x,lo),hi) with minNum/maxNum;
//   smoothstep, step, mod, fract per the GLSL 4.x spec formulas; radians = d * fl(pi/180);
//   matrices are column-major, m[col][row], M*v = c0*v.x + c1*v.y + c2*v.z.
//   sin cos tan exp pow acos atan come from sbx_math.h (bit-compatible with glibc 2.39).
#ifndef SBX_VEC_CUH_
#define SBX_VEC_CUH_

#include "sbx_math.h"

#define SBX_FN __device__ __forceinline__

// FLOAT -> UNORM8 of a D3D11 render target (the reference's 8-bit hosts, include/sbx.h): NaN -> 0, clamp to
// [0, 1], * 255 + 0.5 in fp32, truncate.  fmaxf/fminf return the non-NaN operand, so NaN lands on 0.
__device__ __forceinline__ unsigned sbx_unorm8(float v) {
    return (unsigned)__float2int_rz(__fadd_rn(__fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f), 0.5f));
}
__device__ __forceinline__ unsigned sbx_pack_unorm8(float r, float g, float b, float a) {
    return sbx_unorm8(r) | (sbx_unorm8(g) << 8) | (sbx_unorm8(b) << 16) | (sbx_unorm8(a) << 24);
}


namespace sbx_glsl {

struct vec2; struct vec3; struct vec4;

// read-only swizzle views; they alias the component array of the owning vector
template <int N, int A, int B> struct swz2 {
    float v[N];
    SBX_FN operator vec2() const;
};
template <int N, int A, int B, int C> struct swz3 {
    float v[N];
    SBX_FN operator vec3() const;
};

struct vec2 {
    union {
        float v[2];
        struct { float x, y; };
        struct { float r, g; };
        struct { float s, t; };
        swz2<2, 0, 1> xy; swz2<2, 1, 0> yx; swz2<2, 0, 0> xx; swz2<2, 1, 1> yy;
    };
    SBX_FN vec2() : x(0.0f), y(0.0f) {}
    SBX_FN vec2(float a, float b) : x(a), y(b) {}
    SBX_FN explicit vec2(float a) : x(a), y(a) {}
    SBX_FN float& operator[](int i) { return v[i]; }
    SBX_FN const float& operator[](int i) const { return v[i]; }
};

struct vec3 {
    union {
        float v[3];
        struct { float x, y, z; };
        struct { float r, g, b; };
        swz2<3, 0, 1> xy; swz2<3, 0, 2> xz; swz2<3, 1, 2> yz; swz2<3, 1, 0> yx; swz2<3, 2, 0> zx;
        swz2<3, 2, 1> zy; swz2<3, 0, 0> xx; swz2<3, 1, 1> yy; swz2<3, 2, 2> zz;
        swz3<3, 0, 1, 2> xyz; swz3<3, 0, 2, 2> xzz; swz3<3, 2, 0, 2> zxz; swz3<3, 2, 2, 0> zzx;
        swz3<3, 0, 2, 1> xzy; swz3<3, 1, 0, 2> yxz; swz3<3, 1, 2, 0> yzx; swz3<3, 2, 0, 1> zxy;
        swz3<3, 2, 1, 0> zyx; swz3<3, 0, 1, 2> rgb; swz3<3, 0, 0, 0> xxx; swz3<3, 1, 1, 1> yyy;
        swz3<3, 2, 2, 2> zzz; swz3<3, 0, 1, 1> xyy; swz3<3, 1, 0, 1> yxy; swz3<3, 1, 1, 0> yyx;
    };
    SBX_FN vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    SBX_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    SBX_FN explicit vec3(float a) : x(a), y(a), z(a) {}
    SBX_FN vec3(const vec2& a, float c) : x(a.x), y(a.y), z(c) {}
    SBX_FN vec3(float a, const vec2& b) : x(a), y(b.x), z(b.y) {}
    SBX_FN float& operator[](int i) { return v[i]; }
    SBX_FN const float& operator[](int i) const { return v[i]; }
};

struct vec4 {
    union {
        float v[4];
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<4, 0, 1> xy; swz2<4, 2, 3> zw; swz2<4, 0, 2> xz; swz2<4, 1, 2> yz;
        swz3<4, 0, 1, 2> xyz; swz3<4, 0, 1, 2> rgb;
    };
    SBX_FN vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    SBX_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    SBX_FN explicit vec4(float a) : x(a), y(a), z(a), w(a) {}
    SBX_FN vec4(const vec3& a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
    SBX_FN vec4(const vec2& a, float c, float d) : x(a.x), y(a.y), z(c), w(d) {}
    SBX_FN vec4(const vec2& a, const vec2& b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    SBX_FN float& operator[](int i) { return v[i]; }
    SBX_FN const float& operator[](int i) const { return v[i]; }
};

template <int N, int A, int B> SBX_FN swz2<N, A, B>::operator vec2() const { return vec2(v[A], v[B]); }
template <int N, int A, int B, int C> SBX_FN swz3<N, A, B, C>::operator vec3() const {
    return vec3(v[A], v[B], v[C]);
}

// ---- componentwise operators ------------------------------------------------------------------
SBX_FN vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
SBX_FN vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
SBX_FN vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
SBX_FN vec2 operator/(const vec2& a, const vec2& b) { return vec2(a.x / b.x, a.y / b.y); }
SBX_FN vec2 operator+(const vec2& a, float b) { return vec2(a.x + b, a.y + b); }
SBX_FN vec2 operator-(const vec2& a, float b) { return vec2(a.x - b, a.y - b); }
SBX_FN vec2 operator*(const vec2& a, float b) { return vec2(a.x * b, a.y * b); }
SBX_FN vec2 operator/(const vec2& a, float b) { return vec2(a.x / b, a.y / b); }
SBX_FN vec2 operator+(float a, const vec2& b) { return vec2(a + b.x, a + b.y); }
SBX_FN vec2 operator-(float a, const vec2& b) { return vec2(a - b.x, a - b.y); }
SBX_FN vec2 operator*(float a, const vec2& b) { return vec2(a * b.x, a * b.y); }
SBX_FN vec2 operator/(float a, const vec2& b) { return vec2(a / b.x, a / b.y); }
SBX_FN vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }

SBX_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
SBX_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
SBX_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
SBX_FN vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
SBX_FN vec3 operator+(const vec3& a, float b) { return vec3(a.x + b, a.y + b, a.z + b); }
SBX_FN vec3 operator-(const vec3& a, float b) { return vec3(a.x - b, a.y - b, a.z - b); }
SBX_FN vec3 operator*(const vec3& a, float b) { return vec3(a.x * b, a.y * b, a.z * b); }
SBX_FN vec3 operator/(const vec3& a, float b) { return vec3(a.x / b, a.y / b, a.z / b); }
SBX_FN vec3 operator+(float a, const vec3& b) { return vec3(a + b.x, a + b.y, a + b.z); }
SBX_FN vec3 operator-(float a, const vec3& b) { return vec3(a - b.x, a - b.y, a - b.z); }
SBX_FN vec3 operator*(float a, const vec3& b) { return vec3(a * b.x, a * b.y, a * b.z); }
SBX_FN vec3 operator/(float a, const vec3& b) { return vec3(a / b.x, a / b.y, a / b.z); }
SBX_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }

SBX_FN vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SBX_FN vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
SBX_FN vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
SBX_FN vec4 operator/(const vec4& a, const vec4& b) { return vec4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
SBX_FN vec4 operator+(const vec4& a, float b) { return vec4(a.x + b, a.y + b, a.z + b, a.w + b); }
SBX_FN vec4 operator-(const vec4& a, float b) { return vec4(a.x - b, a.y - b, a.z - b, a.w - b); }
SBX_FN vec4 operator*(const vec4& a, float b) { return vec4(a.x * b, a.y * b, a.z * b, a.w * b); }
SBX_FN vec4 operator/(const vec4& a, float b) { return vec4(a.x / b, a.y / b, a.z / b, a.w / b); }
SBX_FN vec4 operator+(float a, const vec4& b) { return vec4(a + b.x, a + b.y, a + b.z, a + b.w); }
SBX_FN vec4 operator-(float a, const vec4& b) { return vec4(a - b.x, a - b.y, a - b.z, a - b.w); }
SBX_FN vec4 operator*(float a, const vec4& b) { return vec4(a * b.x, a * b.y, a * b.z, a * b.w); }
SBX_FN vec4 operator/(float a, const vec4& b) { return vec4(a / b.x, a / b.y, a / b.z, a / b.w); }
SBX_FN vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

#define SBX_COMPOUND(V)                                                          \
    SBX_FN V& operator+=(V& a, const V& b) { a = a + b; return a; }              \
    SBX_FN V& operator-=(V& a, const V& b) { a = a - b; return a; }              \
    SBX_FN V& operator*=(V& a, const V& b) { a = a * b; return a; }              \
    SBX_FN V& operator/=(V& a, const V& b) { a = a / b; return a; }              \
    SBX_FN V& operator+=(V& a, float b) { a = a + b; return a; }                 \
    SBX_FN V& operator-=(V& a, float b) { a = a - b; return a; }                 \
    SBX_FN V& operator*=(V& a, float b) { a = a * b; return a; }                 \
    SBX_FN V& operator/=(V& a, float b) { a = a / b; return a; }
SBX_COMPOUND(vec2) SBX_COMPOUND(vec3) SBX_COMPOUND(vec4)
#undef SBX_COMPOUND

// ---- packed fp32 pairs (sm_100a FFMA2: one issue slot, two IEEE binary32 operations) ----------
// The march loops are instruction-issue bound (profiles/), and FFMA2 is the only packed fp32
// instruction Blackwell has, so every packed operation is phrased as an fma:
//     a*b   = fma(a, b, -0)      exact: adding -0 changes nothing, including the sign of a zero
//     a+b   = fma(a, 1, b)       a-b = fma(b, -1, a)
// each rounding ONCE, exactly like the scalar operation it replaces.  The constants 1, -0, -1 are
// read from __constant__ memory the compiler cannot see through: ptxas 12.9 contracts
// mul.rn.f32x2 + add.rn.f32x2 (and fma(fma(a,b,-0),1,c)) into ONE FFMA2 even under -fmad=false
// (tools/ubench/f32x2.cu), which would change the rounding; with opaque constants no
// multiply/add pair is visible to it.  pk_fma is a genuine fused multiply-add.
__constant__ float sbx_pk_const[4] = {1.0f, -0.0f, -1.0f, 0.0f};
SBX_FN float2 pk(float a, float b) { return make_float2(a, b); }
SBX_FN float2 pk(float a) { return make_float2(a, a); }
SBX_FN float2 pk_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
SBX_FN float2 pk_mul(float2 a, float2 b) { return __ffma2_rn(a, b, pk(sbx_pk_const[1])); }
SBX_FN float2 pk_mul(float2 a, float b) { return __ffma2_rn(a, pk(b), pk(sbx_pk_const[1])); }
SBX_FN float2 pk_add(float2 a, float2 b) { return __ffma2_rn(a, pk(sbx_pk_const[0]), b); }
SBX_FN float2 pk_sub(float2 a, float2 b) { return __ffma2_rn(b, pk(sbx_pk_const[2]), a); }
SBX_FN float2 pk_one_minus(float2 a) { return __ffma2_rn(a, pk(sbx_pk_const[2]), pk(1.0f)); }
// x*(1-a) + y*a per lane, (1-a) given: the body of mix() (three roundings per lane)
SBX_FN float2 pk_mix(float2 x, float2 y, float one_minus_a, float a) { return pk_add(pk_mul(x, one_minus_a), pk_mul(y, a)); }
SBX_FN float2 pk_mix(float2 x, float2 y, float2 one_minus_a, float2 a) { return pk_add(pk_mul(x, one_minus_a), pk_mul(y, a)); }

// ---- scalar builtins --------------------------------------------------------------------------
// -DSBX_MATH_OUTLINE: one out-of-line copy of each transcendental per kernel instead of one per call site.  For
// apps whose scene function is inlined many times (APP_VINYL: 13 688 SASS instructions, 219 KB, instruction-cache
// bound) the call costs less than the misses; apps with a transcendental in their hot loop keep them inline.
#ifdef SBX_MATH_OUTLINE
#define SBX_TRANSCENDENTAL static __device__ __noinline__
#else
#define SBX_TRANSCENDENTAL SBX_FN
#endif
SBX_TRANSCENDENTAL float sin(float a) { return sbx_sinf(a); }
SBX_TRANSCENDENTAL float cos(float a) { return sbx_cosf(a); }
// both at once from one argument reduction; each equals sin(a) / cos(a) bit for bit (sbx_math.h, sbx_sincosf)
struct sbx_sincos_t { float s, c; };
SBX_TRANSCENDENTAL sbx_sincos_t sbx_sincos(float a) { sbx_sincos_t r; sbx_sincosf(a, &r.s, &r.c); return r; }
SBX_TRANSCENDENTAL float tan(float a) { return sbx_tanf(a); }
SBX_TRANSCENDENTAL float exp(float a) { return sbx_expf(a); }
SBX_TRANSCENDENTAL float pow(float a, float b) { return sbx_powf(a, b); }
SBX_TRANSCENDENTAL float acos(float a) { return sbx_acosf(a); }
SBX_TRANSCENDENTAL float atan(float a) { return sbx_atanf(a); }
SBX_TRANSCENDENTAL float atan(float y, float x) { return sbx_atan2f(y, x); }
SBX_FN float sqrt(float a) { return __fsqrt_rn(a); }
SBX_FN float abs(float a) { return fabsf(a); }
SBX_FN float floor(float a) { return floorf(a); }
SBX_FN float ceil(float a) { return ceilf(a); }
SBX_FN float min(float a, float b) { return fminf(a, b); }      // IEEE minNum
SBX_FN float max(float a, float b) { return fmaxf(a, b); }      // IEEE maxNum
SBX_FN float fract(float a) { return a - floorf(a); }
SBX_FN float mod(float a, float b) { return a - b * floorf(a / b); }
SBX_FN float clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
SBX_FN float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
SBX_FN float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
SBX_FN float smoothstep(float e0, float e1, float x) {
    const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
SBX_FN float radians(float d) { return d * 0.017453292519943295f; }
SBX_FN float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }

// ---- vector builtins --------------------------------------------------------------------------
#define SBX_MAP1(F)                                                                      \
    SBX_FN vec2 F(const vec2& a) { return vec2(F(a.x), F(a.y)); }                        \
    SBX_FN vec3 F(const vec3& a) { return vec3(F(a.x), F(a.y), F(a.z)); }                \
    SBX_FN vec4 F(const vec4& a) { return vec4(F(a.x), F(a.y), F(a.z), F(a.w)); }
SBX_MAP1(sin) SBX_MAP1(cos) SBX_MAP1(exp) SBX_MAP1(sqrt) SBX_MAP1(abs) SBX_MAP1(floor) SBX_MAP1(fract)
#undef SBX_MAP1
#define SBX_VFUNCS(V, ...)                                                               \
    SBX_FN V min(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = min(a.v[i], b.v[i]); return r; } \
    SBX_FN V max(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = max(a.v[i], b.v[i]); return r; } \
    SBX_FN V min(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = min(a.v[i], b); return r; }        \
    SBX_FN V max(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = max(a.v[i], b); return r; }        \
    SBX_FN V mod(const V& a, float b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = mod(a.v[i], b); return r; }        \
    SBX_FN V mod(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = mod(a.v[i], b.v[i]); return r; } \
    SBX_FN V pow(const V& a, const V& b) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = pow(a.v[i], b.v[i]); return r; } \
    SBX_FN V clamp(const V& a, float lo, float hi) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = clamp(a.v[i], lo, hi); return r; } \
    SBX_FN V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = mix(a.v[i], b.v[i], t); return r; } \
    SBX_FN V mix(const V& a, const V& b, const V& t) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = mix(a.v[i], b.v[i], t.v[i]); return r; } \
    SBX_FN V step(float e, const V& a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = step(e, a.v[i]); return r; }      \
    SBX_FN V smoothstep(float e0, float e1, const V& a) { V r; for (int i = 0; i < __VA_ARGS__; ++i) r.v[i] = smoothstep(e0, e1, a.v[i]); return r; }
SBX_VFUNCS(vec2, 2) SBX_VFUNCS(vec3, 3) SBX_VFUNCS(vec4, 4)
#undef SBX_VFUNCS

SBX_FN float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
SBX_FN float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SBX_FN float dot(const vec4& a, const vec4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
SBX_FN float length(const vec2& a) { return __fsqrt_rn(dot(a, a)); }
SBX_FN float length(const vec3& a) { return __fsqrt_rn(dot(a, a)); }
SBX_FN float length(const vec4& a) { return __fsqrt_rn(dot(a, a)); }
SBX_FN vec2 normalize(const vec2& a) { return a / length(a); }
SBX_FN vec3 normalize(const vec3& a) { return a / length(a); }
SBX_FN vec4 normalize(const vec4& a) { return a / length(a); }
SBX_FN float distance(const vec3& a, const vec3& b) { return length(a - b); }
SBX_FN vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// ---- matrices ---------------------------------------------------------------------------------
struct mat2 {
    vec2 c[2];
    SBX_FN mat2() {}
    SBX_FN mat2(float a, float b, float d, float e) { c[0] = vec2(a, b); c[1] = vec2(d, e); }
    SBX_FN vec2& operator[](int i) { return c[i]; }
    SBX_FN const vec2& operator[](int i) const { return c[i]; }
};
struct mat3 {
    vec3 c[3];
    SBX_FN mat3() {}
    SBX_FN mat3(float a, float b, float cc, float d, float e, float f, float g, float h, float i) {
        c[0] = vec3(a, b, cc); c[1] = vec3(d, e, f); c[2] = vec3(g, h, i);
    }
    SBX_FN mat3(const vec3& a, const vec3& b, const vec3& cc) { c[0] = a; c[1] = b; c[2] = cc; }
    SBX_FN vec3& operator[](int i) { return c[i]; }
    SBX_FN const vec3& operator[](int i) const { return c[i]; }
};
SBX_FN vec2 operator*(const mat2& m, const vec2& v) { return m.c[0] * v.x + m.c[1] * v.y; }
SBX_FN vec3 operator*(const mat3& m, const vec3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
SBX_FN vec2 operator*(const vec2& v, const mat2& m) { return vec2(dot(v, m.c[0]), dot(v, m.c[1])); }
SBX_FN vec3 operator*(const vec3& v, const mat3& m) { return vec3(dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2])); }
SBX_FN mat3 operator*(const mat3& a, const mat3& b) { return mat3(a * b.c[0], a * b.c[1], a * b.c[2]); }
SBX_FN mat2 operator*(const mat2& a, const mat2& b) { mat2 r; r.c[0] = a * b.c[0]; r.c[1] = a * b.c[1]; return r; }

// GLSL reflect for app headers that do not include util_optics.h (src/app_vinyl.h:316); util_optics.h's own
// member definition (src/util_optics.h:16-22) shadows this one where it is included.  Same expression.
SBX_FN vec3 reflect(const vec3& incident, const vec3& normal) { return incident - 2.0f * dot(normal, incident) * normal; }

}  // namespace sbx_glsl
#endif  // SBX_VEC_CUH_
