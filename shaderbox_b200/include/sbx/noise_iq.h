// noise_iq.h -- 3-D value noise (replaces src/noise_iq.h:5-29; the `#if 1` arithmetic branch).
//
// The reference hashes a lattice index with fract(sin(n)*753.5453123) -- 8 sines per noise call,
// >95 % of APP_CLOUDS' work.  The lattice index n = px + 157 py + 113 pz is
// always an INTEGER-valued float (floor() results, integer weights; rounding an integer to fp32
// still yields an integer), so hash(n) is memoised.  The table is laid out for noise_iq's access
// pattern AND for the packed arithmetic below: entry k (32 bytes, one L2 sector) holds the eight
// corners of the cell whose base index is n = hash_lo + k, z-neighbours adjacent:
//     float4 lo = { h(n),     h(n+113), h(n+1),   h(n+114) }     (x0y0z0, x0y0z1, x1y0z0, x1y0z1)
//     float4 hi = { h(n+157), h(n+270), h(n+158), h(n+271) }     (x0y1z0, x0y1z1, x1y1z0, x1y1z1)
// so the corners arrive with two adjacent 16-byte read-only loads, already paired as (z0, z1)
// lanes: the x and y interpolations run on both z-slices at once with FFMA2 (sbx_vec.cuh, pk_*).
// It is filled on the device by sbx_hash_table_kernel with this very file's arithmetic; lattice
// indices outside the table take the arithmetic path, so results are bit-identical either way.

// arithmetic definition (src/noise_iq.h:5-9); out of line: it is the rare path once the memo
// table is in place, and eight inlined copies per noise call would only bloat the hot loop
static __device__ __noinline__ float sbx_hash_arith(float n) { return fract(sin(n) * 753.5453123f); }

SBX_FN float hash(_in(float) n) {
    // (n + 1.5*2^23) - 1.5*2^23 == n  <=>  n is an integer with |n| < 2^22
    const float shifted = n + 12582912.0f;
    const unsigned k = (unsigned)(__float_as_int(shifted) - sbx_L->hash_bias);
    if (k < (unsigned)sbx_L->hash_len && (shifted - 12582912.0f) == n) return __ldg(&sbx_L->hash_tab[2u * k].x);
    return sbx_hash_arith(n);
}

// the two bilinear z-slices of the cell at lattice position p, weights (wx, wy) already smoothed:
//   .x = mix(mix(h000,h100,wx), mix(h010,h110,wx), wy)      .y = the same on the z+1 slice
SBX_FN float2 sbx_noise_slices(float px, float py, float pz, float2 wxy) {
    const float n = px + py * 157.0f + 113.0f * pz;      // lattice index, stride (1, 157, 113)
    const unsigned k = (unsigned)(__float_as_int(n + 12582912.0f) - sbx_L->hash_bias);
    float2 c00, c10, c01, c11;                           // (z0, z1) pairs of the corners x?y?
    if (k < (unsigned)sbx_L->hash_span) {
        const float4 lo = __ldg(sbx_L->hash_tab + 2u * k), hi = __ldg(sbx_L->hash_tab + 2u * k + 1u);
        c00 = pk(lo.x, lo.y); c10 = pk(lo.z, lo.w); c01 = pk(hi.x, hi.y); c11 = pk(hi.z, hi.w);
    } else {
        c00 = pk(sbx_hash_arith(n + 0.0f), sbx_hash_arith(n + 113.0f));
        c10 = pk(sbx_hash_arith(n + 1.0f), sbx_hash_arith(n + 114.0f));
        c01 = pk(sbx_hash_arith(n + 157.0f), sbx_hash_arith(n + 270.0f));
        c11 = pk(sbx_hash_arith(n + 158.0f), sbx_hash_arith(n + 271.0f));
    }
    const float2 axy = pk_one_minus(wxy);
    return pk_mix(pk_mix(c00, c10, axy.x, wxy.x), pk_mix(c01, c11, axy.x, wxy.x), axy.y, wxy.y);
}

// smoothstep weights f*f*(3 - 2f) of src/noise_iq.h:16 on a pair: 2f is exact, so fma(f, -2, 3)
// rounds exactly like 3 - 2f
SBX_FN float2 sbx_noise_weight(float2 f) { return pk_mul(pk_mul(f, f), pk_fma(f, pk(-2.0f), pk(3.0f))); }
SBX_FN float sbx_noise_weight(float f) { return f * f * __fmaf_rn(f, -2.0f, 3.0f); }
// mix(y0, y1, wz) of the two slices
SBX_FN float sbx_noise_zmix(float2 ys, float wz) {
    const float2 z = pk_mul(ys, pk(1.0f - wz, wz));
    return z.x + z.y;
}

SBX_FN float noise_iq(_in(vec3) x) {
    const float px = floor(x.x), py = floor(x.y), pz = floor(x.z);
    const float2 wxy = sbx_noise_weight(pk_sub(pk(x.x, x.y), pk(px, py)));      // fract(x) = x - floor(x)
    const float wz = sbx_noise_weight(x.z - pz);
    return sbx_noise_zmix(sbx_noise_slices(px, py, pz, wxy), wz);
}
