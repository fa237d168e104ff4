// noise_iq.h -- 3-D value noise (replaces src/noise_iq.h:5-29; the `#if 1` arithmetic branch).
//
// The reference hashes a lattice index with fract(sin(n)*753.5453123) -- 8 sines per noise call,
// >95 % of APP_CLOUDS' work.  The lattice index n = px + 157 py + 113 pz is
// always an INTEGER-valued float (floor() results, integer weights; rounding an integer to fp32
// still yields an integer), so hash(n) is memoised.  The table is laid out for the access pattern
// of noise_iq: entry k is the float4 { h(n), h(n+1), h(n+157), h(n+158) } with n = hash_lo + k --
// one z-slice of the cell -- so the eight corners are TWO 16-byte read-only loads (entries k and
// k+113) instead of eight 4-byte ones.  It is filled on the device by sbx_hash_table_kernel with
// this very file's arithmetic; lattice indices outside the table take the arithmetic path, so
// results are bit-identical either way.

// arithmetic definition (src/noise_iq.h:5-9); out of line: it is the rare path once the memo
// table is in place, and eight inlined copies per noise call would only bloat the hot loop
static __device__ __noinline__ float sbx_hash_arith(float n) { return fract(sin(n) * 753.5453123f); }

SBX_FN float hash(_in(float) n) {
    // (n + 1.5*2^23) - 1.5*2^23 == n  <=>  n is an integer with |n| < 2^22
    const float shifted = n + 12582912.0f;
    const unsigned k = (unsigned)(__float_as_int(shifted) - sbx_L->hash_bias);
    if (k < (unsigned)sbx_L->hash_len && (shifted - 12582912.0f) == n) return __ldg(&sbx_L->hash_tab[k].x);
    return sbx_hash_arith(n);
}

SBX_FN float noise_iq(_in(vec3) x) {
    const vec3 p = floor(x);
    vec3 f = fract(x);
    // smoothstep weights f*f*(3 - 2f): 2f is exact, so fma(f, -2, 3) rounds exactly like 3 - 2f
    f = vec3(f.x * f.x * __fmaf_rn(f.x, -2.0f, 3.0f), f.y * f.y * __fmaf_rn(f.y, -2.0f, 3.0f),
             f.z * f.z * __fmaf_rn(f.z, -2.0f, 3.0f));

    const float n = p.x + p.y * 157.0f + 113.0f * p.z;   // lattice index, stride (1, 157, 113)
    float h000, h100, h010, h110, h001, h101, h011, h111;
    const unsigned k = (unsigned)(__float_as_int(n + 12582912.0f) - sbx_L->hash_bias);
    if (k < (unsigned)sbx_L->hash_span) {                // entries k and k+113 are tabulated
        const float4 z0 = __ldg(sbx_L->hash_tab + k), z1 = __ldg(sbx_L->hash_tab + k + 113);
        h000 = z0.x; h100 = z0.y; h010 = z0.z; h110 = z0.w;
        h001 = z1.x; h101 = z1.y; h011 = z1.z; h111 = z1.w;
    } else {
        h000 = sbx_hash_arith(n + 0.0f);   h100 = sbx_hash_arith(n + 1.0f);
        h010 = sbx_hash_arith(n + 157.0f); h110 = sbx_hash_arith(n + 158.0f);
        h001 = sbx_hash_arith(n + 113.0f); h101 = sbx_hash_arith(n + 114.0f);
        h011 = sbx_hash_arith(n + 270.0f); h111 = sbx_hash_arith(n + 271.0f);
    }
    return mix(mix(mix(h000, h100, f.x), mix(h010, h110, f.x), f.y),
               mix(mix(h001, h101, f.x), mix(h011, h111, f.x), f.y), f.z);
}
