/* sbx.h -- C ABI of the B200-native shaderbox pixel path.
 *
 * What this boundary replaces.  shaderbox has no FFI: its boundary is a SOURCE-level plugin
 * contract.  An app header (src/app_*.h) defines setup_camera / setup_scene / render / FOV and
 * then includes src/main.h, which supplies
 *
 *     void mainImage(out vec4 fragColor, in vec2 fragCoord)            (src/main.h:6-53)
 *
 * and a HOST calls that once per pixel after providing the uniforms u_res / u_time / u_mouse
 * (src/uniform_buffer.h:26-36) plus the optional aux block (src/uniform_buffer.h:39-60).
 * The hosts in the reference are VML's SDL_app.cpp (C++ CPU; src/Makefile:21, not in the tree),
 * util/hlsltoy (D3D11; util/hlsltoy/src/hlsltoy.cpp:382-426 compiles the shader at run time and
 * uploads the two cbuffers) and shadertoy.com.  This library IS such a host: the per-pixel loop
 * becomes one CUDA grid launch on sm_100a, and the frame is written as raw RGBA32F.
 *
 * Conventions
 *   - every entry point returns 0 (SBX_OK) or a negative sbx_status; no exceptions cross the ABI
 *   - frames are row-major float4 (R,G,B,A=1), row 0 is fragCoord.y = 0.5 (GL / shadertoy
 *     bottom-up convention, src/main.h:12), pixel centre fragCoord = (x+0.5, y+0.5)
 *   - a context is bound to ONE device; calls on one context are serialised by the caller, but launches may go to
 *     any streams: per-context device state (math tables, the lattice-hash memo) is complete before the call that
 *     builds it returns, and per-launch state (sequence times, completion counters) is per launch
 *   - there is NO CPU path: if no sm_100-class device / kernel image is available the calls fail
 */
#ifndef SBX_H_
#define SBX_H_

#if !defined(__CUDACC_RTC__)
#include <stddef.h>   /* size_t */
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum sbx_status {
    SBX_OK = 0,
    SBX_ERR_INVALID = -1,      /* bad argument (null pointer, non-positive size, bad shard) */
    SBX_ERR_NO_DEVICE = -2,    /* no usable CUDA device / driver */
    SBX_ERR_UNKNOWN_APP = -3,  /* app name not registered / not loaded */
    SBX_ERR_CUDA = -4,         /* a CUDA runtime / driver call failed (see sbx_last_error) */
    SBX_ERR_COMPILE = -5,      /* run-time compilation of an app header failed (log in sbx_last_error) */
    SBX_ERR_NOMEM = -6,
    SBX_ERR_UNSUPPORTED = -7
} sbx_status;

/* The uniform block.  Field names, meaning and defaults follow src/uniform_buffer.h:
 *   main block  :26-30 (u_res is width/height below, u_mouse, u_time)
 *   APP_CLOUDS  :41-54 ; APP_SDF_AO :57-58
 * u_mouse is the shadertoy iMouse vec4 (src/uniform_buffer.h:35). */
typedef struct sbx_params {
    int   width, height;          /* u_res */
    float u_time;
    float u_mouse[4];
    /* aux_uniform_buffer_t, APP_CLOUDS */
    float wind_dir[3];
    float sun_dir[3];
    float sun_color[3];
    float sun_power;
    int   cld_march_steps;
    int   illum_march_steps;
    float sigma_scattering;
    float cld_coverage;
    float cld_thick;
    float atm_radius;
    float atm_ground_y;
    /* aux_uniform_buffer_t, APP_SDF_AO */
    float fog_density;
    float fog_falloff;
} sbx_params;

/* Which rows of the frame a call renders.  The frame is cut into stripes of `stripe_rows`
 * consecutive rows; stripe s belongs to part (s % n_parts).  A call renders the stripes of `part`
 * and stores them COMPACTED (in increasing row order) so each part's output is one contiguous
 * message for the gather.  {1,1,0} (or a zeroed struct) = the whole frame. */
typedef struct sbx_shard {
    int stripe_rows;
    int n_parts;
    int part;
} sbx_shard;

typedef struct sbx_timing {
    float kernel_ms;     /* device time of the last render kernel (CUDA events on its stream) */
    float h2d_ms;        /* uniform / table upload, if any */
    float d2h_ms;        /* frame read-back, if the call had a host destination */
    int   launches;      /* kernels launched by the last call */
    int   grid_blocks, block_threads, regs_per_thread, blocks_per_sm;
    int   zero_copy;     /* 1 if sbx_render_host stored straight into a pinned+mapped host frame */
    int   lanes_per_pixel; /* 1, or P for a cooperative image (P lanes share one pixel's march) */
    int   tail_rows;       /* rows at the end of the launch marched with tail_lanes_per_pixel lanes (hybrid image), else 0 */
    int   tail_lanes_per_pixel;
} sbx_timing;

typedef struct sbx_ctx sbx_ctx;

/* Fill *p with the reference defaults (src/uniform_buffer.h:41-58), width x height, u_time = 0. */
int sbx_default_params(sbx_params* p, int width, int height);

/* Number of CUDA devices the driver shows (0 without a driver / device); never fails. */
int sbx_device_count(void);

/* Create a context on CUDA device `device` (primary context; interoperates with torch). */
int sbx_create(int device, sbx_ctx** out);
void sbx_destroy(sbx_ctx* ctx);

/* Select the app, by its reference define: "APP_EGG", "APP_CLOUDS", "APP_ATMOSPHERE",
 * "APP_PLANET", "APP_RAYTRACER" (src/Makefile:9 `APP = -DAPP_PLANET`).  `variant` picks the
 * implementation: "native" = the hand-written sm_100a kernel built into this library,
 * "plugin" = the kernel image compiled from an UNCHANGED shaderbox app header (see
 * sbx_compile_app / the prebuilt images next to the library), NULL = native if present. */
int sbx_load_app(sbx_ctx* ctx, const char* app_name, const char* variant);

/* Compile an unchanged shaderbox app header (a file that ends in `#include "main.h"`) against the
 * device operator library with NVRTC for sm_100a, register it under `app_name`, and (if
 * `image_out_path` is not NULL) also write the cubin there.  The equivalent of hlsltoy's
 * D3DCompileFromFile (util/hlsltoy/src/hlsltoy.cpp:382-388).  Needs no GPU. */
int sbx_compile_app(sbx_ctx* ctx_or_null, const char* app_header_path, const char* app_name,
                    const char* image_out_path);

/* Number of rows / bytes a shard produces for a frame of `height` rows. */
int sbx_shard_rows(const sbx_shard* shard, int height);

/* Render into DEVICE memory (dev_rgba: rows*width float4, 16-byte aligned) on `stream`
 * (a cudaStream_t passed as void*, NULL = the legacy default stream).  Asynchronous. */
int sbx_render_device(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                      float* dev_rgba, void* stream);

/* Render to HOST memory (host_rgba: rows*width*4 floats).  Synchronous; this is the call the
 * reference-facing hosts make (one frame in, one frame out).  If host_rgba is pinned and mapped
 * (cudaHostAlloc, a pinned torch tensor) the kernel stores into it directly over PCIe; otherwise the
 * frame is rendered in HBM and copied. */
int sbx_render_host(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                    float* host_rgba);

/* The 8-bit output of the reference's presenting hosts.  hlsltoy renders mainImage's colour (already
 * sRGB-encoded by src/main.h:52 linear_to_srgb) into a DXGI_FORMAT_R8G8B8A8_UNORM swap chain
 * (util/hlsltoy/src/hlsltoy.cpp:192); the FLOAT -> UNORM rule of that target is: NaN -> 0, clamp to [0, 1],
 * multiply by 255, add 0.5, truncate.  These entry points apply it inside the render kernel's store (one
 * 32-bit store per pixel, R in the low byte; a quarter of the frame bytes to HBM / over PCIe).  Frames are
 * rows*width*4 bytes, same row order and sharding as the float entry points; host frames may be pinned+mapped
 * (zero-copy) or pageable. */
int sbx_render_device_rgba8(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                            unsigned char* dev_rgba8, void* stream);
int sbx_render_host_rgba8(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                          unsigned char* host_rgba8);

/* Render the rows of `shard` straight into a FULL frame (height*width float4) at their frame rows --
 * no compaction, no gather, no unshard.  dev_frame may live on ANOTHER GPU of the box (a pointer
 * obtained from sbx_frame_import): the kernel's float4 stores then travel over NVLink while the
 * remaining pixels are still being computed, which fuses the multi-GPU "gather" into the render
 * kernel.  The caller orders completion across GPUs (e.g. a barrier after the launches). */
int sbx_render_frame(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                     float* dev_frame, void* stream);

/* One GPU's part of a full frame, for hosts that split a frame over the GPUs of a box (the pixels are independent --
 * src/main.h:6-53 reads nothing but uniforms -- so any partition is exact):
 *   rows        interleaved row stripes, as above ({1,1,0} or zeroed = every row), and/or
 *   tile_parts  > 1: of the 8-pixel-wide warp tiles of each tile row, this launch renders the columns tx with
 *               (tx + first_row_of_tile / 4) % tile_parts == tile_part -- a checkerboard of tiles.  Every part then
 *               holds the same share of EVERY row, so parts are balanced by construction whatever the scene's
 *               vertical structure (APP_CLOUDS: the sky/horizon/cloud bands).
 *   done_flag   NULL, or a device-visible 32-bit word (this GPU's memory, a peer GPU's -- typically next to the frame
 *               the pixels go to -- or mapped host memory): the launch's last thread block stores done_value there,
 *               with system-scope release ordering, after every pixel of the launch has been stored.  Wait for it
 *               with sbx_stream_wait_flags (stream-ordered, no host involvement) or by reading it.
 * sbx_render_frame_part is sbx_render_frame with this description; asynchronous on `stream`. */
typedef struct sbx_frame_part {
    sbx_shard rows;
    int tile_parts, tile_part;
    unsigned* done_flag;
    unsigned done_value;
} sbx_frame_part;
int sbx_render_frame_part(sbx_ctx* ctx, const sbx_params* p, const sbx_frame_part* part, float* dev_frame, void* stream);
/* The same signal from the stream instead of the kernel: after everything enqueued on `stream` so far has completed,
 * store `value` at dev_flag (cuStreamWriteValue32, preceded by a system-scope fence).  No kernel-side cost; measured on
 * 8 B200s it is the cheaper of the two when the pixels travel over NVLink (the in-kernel flag makes every thread block
 * wait for its peer writes to be acknowledged: +7 % kernel time on a 0.36 ms launch), DESIGN.md 7b. */
int sbx_stream_write_flag(sbx_ctx* ctx, unsigned* dev_flag, unsigned value, void* stream);
/* Make `stream` wait (on the device, cuStreamWaitValue32) until each of the n flags at dev_flags is >= value. */
int sbx_stream_wait_flags(sbx_ctx* ctx, const unsigned* dev_flags, int n, unsigned value, void* stream);

/* A time sequence in ONE launch: frame k is rendered with u_time = times[k] (the animation loop of the
 * reference's hosts -- `iGlobalTime` advancing per presented frame, src/uniform_buffer.h:34 -- batched, so
 * small frames fill the GPU and the launch cost is paid once).  Output: n_frames consecutive frames (each
 * rows*width float4, same row order / sharding as sbx_render_device).  Each frame is bit-identical to the
 * one sbx_render_* produces for that u_time.  `times` is a HOST array of n_frames floats (1 <= n <= 65535); it is
 * consumed before the call returns (copied to a pinned staging slot), and the device entry point is asynchronous. */
int sbx_render_sequence_device(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                               const float* times, int n_frames, float* dev_rgba, void* stream);
int sbx_render_sequence_host(sbx_ctx* ctx, const sbx_params* p, const sbx_shard* shard_or_null,
                             const float* times, int n_frames, float* host_rgba);

/* Frame buffers that can be shared between the per-GPU processes of one box (CUDA IPC):
 * alloc on the owner, export a 64-byte handle, import it in a peer process (enables peer access). */
#define SBX_IPC_HANDLE_BYTES 64
int sbx_frame_alloc(sbx_ctx* ctx, size_t bytes, float** dev_out);
int sbx_frame_free(sbx_ctx* ctx, float* dev);
int sbx_frame_export(sbx_ctx* ctx, const float* dev, unsigned char handle[SBX_IPC_HANDLE_BYTES]);
int sbx_frame_import(sbx_ctx* ctx, const unsigned char handle[SBX_IPC_HANDLE_BYTES], float** dev_out);
int sbx_frame_release(sbx_ctx* ctx, float* imported);
/* A HOST frame every GPU of the box can store into.  `host` is host memory the caller owns -- for one process
 * per GPU, the same POSIX shared-memory mapping in every process.  Register pins and maps it for this
 * context's device and returns the device-side alias; passing that alias to sbx_render_frame makes each
 * rank's render kernel store its stripes straight into the shared host frame over ITS OWN PCIe link, so an
 * N-GPU frame reaches host memory with no gather, no staging copy, and N links in parallel.  The frame is
 * complete once every rank has synchronised its stream. */
int sbx_host_frame_register(sbx_ctx* ctx, void* host, size_t bytes, float** dev_alias_out);
int sbx_host_frame_unregister(sbx_ctx* ctx, void* host);
/* Host frames the render kernel can store into directly: pinned, mapped into every CUDA context of the process
 * (portable).  A frame from sbx_host_alloc passed to sbx_render_host[_rgba8] takes the zero-copy path (the kernel's
 * pixel stores travel over PCIe while other pixels are still being computed); a malloc'd / std::vector frame takes
 * the render-in-HBM + copy path.  This is the allocator a shaderbox host should use for its frame. */
int sbx_host_alloc(sbx_ctx* ctx, size_t bytes, void** host_out);
int sbx_host_free(sbx_ctx* ctx, void* host);
/* Synchronous device->host read of a frame buffer (after ordering `stream`). */
int sbx_frame_read(sbx_ctx* ctx, const float* dev, float* host, size_t bytes, void* stream);

/* ---- one frame over several GPUs of the box, from ONE process -------------------------------------------------
 * The reference's hosts are single programs with one render loop (hlsltoy's message loop,
 * util/hlsltoy/src/hlsltoy.cpp:494-520; VML's SDL_app.cpp): a group gives such a program every GPU of the box
 * behind one call per frame.  devices = NULL means GPUs 0 .. n_gpus-1; a device may be listed more than once (the
 * parts then share that GPU -- useful to exercise the N-part path on a 1-GPU machine).  Each GPU renders its
 * 4-row stripes of the frame (sbx_frame_part) straight into the destination:
 *   sbx_multi_render_device  the group's frame in the FIRST device's memory (peers store over NVLink); asynchronous:
 *                            the frame is complete for work enqueued afterwards on sbx_multi_stream(), or after
 *                            sbx_multi_sync().  *dev_frame_out stays owned by the group.
 *   sbx_multi_render_host    synchronous.  A frame from sbx_host_alloc is written by every GPU over its own PCIe link
 *                            (no gather, no copy); any other host frame is rendered on the device path and copied.
 * Frames are bit-identical to sbx_render_host on one GPU.  Calls on one group are serialised by the caller (like calls on
 * one context).  sbx_multi_ctx(i) exposes the i-th per-GPU context
 * (options, timing).  sbx_multi_last_timing: kernel_ms of each GPU's last launch (after a sync / host render). */
typedef struct sbx_multi sbx_multi;
int sbx_multi_create(const int* devices_or_null, int n_gpus, sbx_multi** out);
void sbx_multi_destroy(sbx_multi* m);
int sbx_multi_gpus(sbx_multi* m);
sbx_ctx* sbx_multi_ctx(sbx_multi* m, int i);
int sbx_multi_load_app(sbx_multi* m, const char* app_name, const char* variant);
int sbx_multi_set_option(sbx_multi* m, const char* key, int value);
int sbx_multi_render_device(sbx_multi* m, const sbx_params* p, float** dev_frame_out);
int sbx_multi_render_host(sbx_multi* m, const sbx_params* p, float* host_rgba);
void* sbx_multi_stream(sbx_multi* m);
int sbx_multi_sync(sbx_multi* m);
int sbx_multi_last_timing(sbx_multi* m, float* kernel_ms_per_gpu, int capacity);
const char* sbx_multi_last_error(sbx_multi* m);

/* Scatter a compacted shard (as produced above / received from a peer) into a full frame on the
 * device: the de-interleave step after the multi-GPU gather. */
int sbx_unshard_device(sbx_ctx* ctx, int width, int height, const sbx_shard* shard,
                       const float* dev_part, float* dev_frame, void* stream);

/* The 3-D noise texture of the reference's USE_NOISE_TEX path (src/app_clouds.h:51-55), baked by
 * util/ddsvolgen/src/ddsvolgen.cpp:52-61,101-116: a size^3 volume of R32G32B32A32_FLOAT voxels, x fastest, voxel
 * (x,y,z) = (fbm_worley_tile((vec3(x,y,z) + .5) / size, 2., 1., .5), 0, 0, 0) with the 4-octave tiled Worley fbm
 * of src/noise_worley.h + src/fbm.h:8.  Renders slices [z0, z0 + nz) into nz*size*size float4 (device or host).
 * sbx_dds_volume_header fills the 148-byte DDS + DX10 header ddsvolgen writes in front of the data
 * (ddsvolgen.cpp:69-92) for a size^3 RGBA32F volume; returns the number of bytes written (148) or < 0. */
int sbx_bake_noise_volume_device(sbx_ctx* ctx, int size, int z0, int nz, float* dev_rgba, void* stream);
int sbx_bake_noise_volume_host(sbx_ctx* ctx, int size, int z0, int nz, float* host_rgba);
int sbx_dds_volume_header(int size, unsigned char* out, int capacity);

/* The two 3-D noise textures of the USE_NOISE_TEX cloud path (`Texture3D u_tex_noise, u_tex_noise_2`,
 * src/app_clouds.h:51-55, bound by hlsltoy from its DDS arguments, util/hlsltoy/src/hlsltoy.cpp:225-239), for the app
 * "APP_CLOUDS_TEX": that branch of app_clouds.h with `SampleLevel(u_sampler0, pos, 0).r` (:69, :77) evaluated by a
 * software sampler -- D3D11 linear filtering, WRAP addressing (hlsltoy.cpp:244-249), 8-bit sub-texel weights; the rule is
 * written down in oracle/sbx_oracle.c and is this library's definition (the reference has no C++ statement of it).
 * Volumes are size^3 R32G32B32A32_FLOAT, x fastest, as sbx_bake_noise_volume_host / ddsvolgen produce; only .r is read.
 * The library keeps its own device copies (single channel, wrapped apron, TMA descriptors); the host arrays are consumed
 * before the call returns. */
int sbx_set_noise_volumes(sbx_ctx* ctx, const float* host_rgba_a, const float* host_rgba_b, int size);

/* Options: "tail_waves_x100" T (default 0 = off) / "tail_max_waves_x100" M: with the default variant of an app that
 * ships a hybrid image, a launch smaller than M/100 waves of resident warps marches its last T/100 waves' worth of
 * rows with 4 lanes per pixel and the rest with one (measured on B200: it does not beat the whole-launch choice
 * below, DESIGN.md 7b -- the long warps are the ones just above the horizon, not the last ones).
 * "trivial_rows_last" 0|1 (default 1): issue the rows a kernel image declares trivial (APP_CLOUDS: the sky under the
 * horizon) at the END of the launch, where they fill the drain instead of delaying the long rays.
 * "record_events" 0|1 (default 1): bracket render launches with the two timing events behind sbx_last_timing.kernel_ms;
 * 0 drops them for sbx_render_frame / sbx_render_frame_part (about 2 us of stream time each; kernel_ms then reads 0).
 * "use_hash_table" 0|1, "hash_table_log2" 9..22 (noise_iq memo table), "host_zero_copy" 0|1,
 * "coop_waves_x100" W (default 250): with the default variant, launches smaller than W/100 waves of resident warps
 * use the app's 4-lanes-per-pixel cooperative image, launches smaller than 2W/100 waves the 2-lane one, if
 * shipped; 0 = never. */
int sbx_set_option(sbx_ctx* ctx, const char* key, int value);

int sbx_last_timing(sbx_ctx* ctx, sbx_timing* out);
/* Profiling hook.  Kernel images built with -DSBX_TRACE (csrc/Makefile `make trace`, variants "native_trace" /
 * "coop_trace" / "hybrid_trace"; not shipped by default) record, per warp of a launch, 4 x 64-bit words
 * { start ns, end ns, SM id, region } at dev_records[4 * warp] (%globaltimer).  The buffer must hold
 * 4 * (grid_blocks * block_threads / 32) words; NULL switches recording off.  Ordinary images ignore it. */
int sbx_set_trace_buffer(sbx_ctx* ctx, unsigned long long* dev_records);
const char* sbx_last_error(sbx_ctx* ctx_or_null);
const char* sbx_strerror(int status);
const char* sbx_version(void);

/* Operator known-answer entry: evaluate one operator of the device library (sdf.h / noise_*.h /
 * fbm.h / volumetric.h / light.h ... ) on n inputs ON THE DEVICE.  `op` names the operator
 * ("noise_iq", "hash", "sd_torus", "sinf", ...); in/out are host arrays of n*in_stride /
 * n*out_stride floats.  Test hook for the operator surface, not a production path. */
int sbx_eval_op(sbx_ctx* ctx, const char* op, const float* in, int in_stride,
                float* out, int out_stride, int n);

#ifdef __cplusplus
}
#endif
#endif /* SBX_H_ */
