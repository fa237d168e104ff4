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

/* A run of consecutive warps of one launch that share a tile shape.  Warp w of the region renders tile
 * (row w / tiles_per_row, column slot w % tiles_per_row); the tile is tile_w x tile_h pixels (8 x 4 with one lane
 * per pixel, (32/P) x 1 with P lanes per pixel) and its first local row is row0 + (w / tiles_per_row) * tile_h. */
typedef struct sbx_region {
    int warps;                  /* tile rows * tiles_per_row */
    int tiles_per_row;          /* column slots per tile row (this part's share of the tile columns, rounded up) */
    int row0, rows;             /* local rows [row0, row0 + rows) */
    int tile_rows;              /* warps / tiles_per_row */
    int first_tile_row;         /* tile rows are ISSUED starting here and wrapping round (the rows below it go last) */
    unsigned long long magic;   /* ceil(2^40 / tiles_per_row) when warps * tiles_per_row < 2^40, else 0 (kernel divides) */
} sbx_region;

typedef struct sbx_launch {
    sbx_params p;               /* uniforms (u_res/u_time/u_mouse + aux block) */
    /* rows: local row lr of this launch is frame row ((lr / stripe) * parts + part) * stripe + lr % stripe */
    int stripe_rows, n_parts, part;
    int local_rows;             /* rows rendered by this launch */
    /* tile-column interleave (frame outputs only): of the tiles of tile row r this launch renders the columns
       tx with (tx + r') % col_parts == col_part, r' = first local row of the tile / 4; col_parts <= 1: every column */
    int col_parts, col_part;
    /* reg[0]: the image's own tile shape (one lane per pixel, or SBX_LANES_PER_PIXEL lanes);  reg[1]: hybrid images
       only -- the rows above reg[0]'s, marched by SBX_HYBRID_LANES lanes per pixel (the tail of the launch) */
    sbx_region reg[2];
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
    const float* times;         /* NULL, or u_time of frame blockIdx.y of a sequence launch (device memory) */
    const void* lut;            /* SBX_LUT_MATH_BYTES of exp2/log2 tables in global memory (sbx_math.h) */
    /* completion signal (NULL = none): every CTA counts itself on done_counter (this device) after its stores;
       the last one resets the counter and stores done_value to done_flag with system-scope release -- done_flag
       may live on a peer GPU (next to the frame the stores went to) or in mapped host memory */
    unsigned* done_counter;
    unsigned* done_flag;
    unsigned done_value;
    /* profiling hook (images built with -DSBX_TRACE only; sbx_set_trace_buffer): warp w of the launch records
       { start ns, end ns, SM id, region } (4 x u64, %globaltimer) at trace[4 * w] */
    unsigned long long* trace;
} sbx_launch;

/* Second kernel parameter of images built with -DSBX_USES_NOISE_TEX (the USE_NOISE_TEX cloud path, src/app_clouds.h:8-9,
 * 51-55): the two 3-D noise textures.  Device layout (sbx_set_noise_volumes): the .r channel only, fp32, with a one-texel
 * apron on every side that holds the WRAPPED neighbours (padded[k] = texel[(k - 1) mod size], k = 0 .. size + 1), so the
 * sampler never wraps an index; rows are `pitch_x` floats long (a multiple of 4: TMA wants 16-byte strides).
 * map[t] is the TMA descriptor of texture t (cuTensorMapEncodeTiled: rank 3, fp32, box 8 x 4 x 4, zero fill outside). */
typedef struct sbx_tex_params {
#if defined(__cplusplus)
    alignas(64)
#endif
    unsigned long long map[2][16];   /* two CUtensorMap (128 bytes each, 64-byte aligned), read from the kernel's parameter space */
    const float* vol[2];             /* padded volumes in global memory */
    int size;                        /* texels per axis (N); the padded volume is (N + 2)^3 */
    int pitch_x;                     /* floats per padded row */
    int pitch_xy;                    /* floats per padded plane = pitch_x * (N + 2) */
    int reserved;
} sbx_tex_params;

#endif
