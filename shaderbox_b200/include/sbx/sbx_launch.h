// sbx_launch.h -- the kernel parameter block shared by the host launcher and every render kernel.
#ifndef SBX_LAUNCH_H_
#define SBX_LAUNCH_H_
#include "sbx.h"
#if !defined(__CUDACC__) && !defined(__CUDACC_RTC__)
#include <vector_types.h>   /* float4 for the host launcher */
#endif

#define SBX_TILE_W 8          /* a warp renders an 8 x 4 pixel tile */
#define SBX_TILE_H 4
#define SBX_HASH_MAGIC_BITS 0x4B400000   /* bits of 1.5*2^23: float(n) + 1.5*2^23 has n in its low mantissa */

typedef struct sbx_launch {
    sbx_params p;               /* uniforms (u_res/u_time/u_mouse + aux block) */
    /* rows: local row lr of this launch is frame row ((lr / stripe) * parts + part) * stripe + lr % stripe */
    int stripe_rows, n_parts, part;
    int local_rows;             /* rows rendered by this launch */
    int tiles_x, tiles_y;       /* warp tiles covering width x local_rows */
    float* out;                 /* local_rows * width float4, compacted */
    /* memoised lattice hash (noise_iq.h): entry k = 2 float4 = the 8 corners of the noise_iq cell with base
       index n = hash_lo + k:  { h(n), h(n+113), h(n+1), h(n+114) }, { h(n+157), h(n+270), h(n+158), h(n+271) } */
    const float4* hash_tab;
    int hash_bias;              /* SBX_HASH_MAGIC_BITS + hash_lo */
    int hash_len;               /* entries */
    int hash_span;              /* == hash_len (every entry is a self-contained cell) */
    int out_is_frame;           /* 0: out rows are this launch's local (compacted) rows; 1: out is the FULL frame
                                   (possibly a peer GPU's, mapped over NVLink) and rows land at their frame row */
    int out_rgba8;              /* 0: out is float4 RGBA32F per pixel; 1: out is one packed R8G8B8A8_UNORM word per pixel */
    unsigned long long tiles_x_magic;   /* ceil(2^40 / tiles_x) when total_tiles * tiles_x < 2^40, else 0 (kernel divides) */
    const float* times;         /* NULL, or u_time of frame blockIdx.y of a sequence launch (device memory) */
    const void* lut;            /* SBX_LUT_MATH_BYTES of exp2/log2 tables in global memory (sbx_math.h) */
} sbx_launch;

#endif
