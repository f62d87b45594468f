/* vr.h — C-ABI of the B200-native volume path tracer (libvr.so).
 *
 * This is the layer that replaces cl-volume-renderer's `opencl_wrapper` (clw_context / clw_vector /
 * clw_image / clw_function, opencl_wrapper/include/clw_*.hpp) for the renderer hot path.  Every entry
 * point cites the reference call site(s) it replaces; paths are relative to the reference checkout.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes only; no C++ types, no exceptions, no torch types.
 *   - every function returning `int` returns VR_OK (0) or a negative vr_status; the message of the
 *     last failure on the calling thread is vr_last_error().  (The reference prints and exit(1)s:
 *     opencl_wrapper/include/clw_helper.hpp:293-309 — the C++ shim in host/ keeps that behaviour.)
 *   - host pointers are caller-owned; device memory is owned by the library behind the handles.
 *   - volumes are `short`, x fastest: voxel (x,y,z) at x + nx*(y + ny*z)   (app/volume_block.hpp:6-34,
 *     app/nrrd_loader.cpp:126-151).  Frames and env maps are RGBA8, row-major.
 *   - one CUDA stream per context; calls are not re-entrant on one context (same as the reference's
 *     single in-order queue, opencl_wrapper/src/clw_context.cpp:38-48).
 *   - there is NO CPU fallback: without a usable CUDA device vr_ctx_create fails with VR_ERR_CUDA.
 */
#ifndef VR_H
#define VR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vr_status {
  VR_OK = 0,
  VR_ERR_INVALID = -1, /* bad argument / bad state */
  VR_ERR_CUDA = -2,    /* CUDA runtime failure (message carries the CUDA error string) */
  VR_ERR_PARSE = -3,   /* transfer-function source not in one of the generated forms */
  VR_ERR_NOMEM = -4
} vr_status;

typedef struct vr_ctx vr_ctx;
typedef struct vr_volume vr_volume;
typedef struct vr_envmap vr_envmap;
typedef struct vr_sdf vr_sdf;
typedef struct vr_renderer vr_renderer;

/* One clause of the run-time generated `is_event_gen` (app/ui.cpp:160-168 + app/tf_part.cpp:55-79).
 * Clauses are tested in order, first match wins.
 *   flags & VR_TF_USE_GRADIENT : the `&& gradient > min_g && gradient < max_g` part is present
 *   flags & VR_TF_THRESHOLD    : the test/bench form `return (value > min_v);` (tests/sdf/sdf_test.cpp:22,
 *                                app/sdf_benchmark.cpp:18) — terminal, does not write a colour
 * rgba = the `int4 tmp_color` of the clause, each (int)(c*255). */
enum { VR_TF_USE_GRADIENT = 1, VR_TF_THRESHOLD = 2 };
typedef struct vr_tf_rect {
  float min_v, max_v, min_g, max_g;
  int32_t flags;
  int32_t rgba[4];
} vr_tf_rect;
#define VR_TF_MAX_RECTS 16

const char* vr_last_error(void);

/* ---- context: replaces clw_context (clw_context.hpp:5-28; main.cpp:13, sdf_test.cpp:14) ------------ */
int vr_ctx_create(int device_ordinal, vr_ctx** out);
void vr_ctx_destroy(vr_ctx* ctx);
int vr_ctx_synchronize(vr_ctx* ctx);
/* cudaStream_t of the context (for CUDA-event timing by the caller). */
void* vr_ctx_stream(vr_ctx* ctx);
/* number of kernels this context has launched so far */
uint64_t vr_ctx_launch_count(const vr_ctx* ctx);
/* number of 3-D CUDA arrays (SDF surfaces, hw-linear step fields) the context holds, in use or cached for reuse */
int vr_ctx_array_count(const vr_ctx* ctx);

/* ---- transfer function ---------------------------------------------------------------------------- */
/* Strict parser for the two generated forms of `is_event_gen` (host only, needs no GPU).
 * Writes at most `cap` clauses, *n = number of clauses found. */
int vr_tf_parse(const char* is_event_gen_src, vr_tf_rect* out, int cap, int* n);
/* Inverse: renders clauses as the text app/ui.cpp:160-168 would have generated (for round-trip tests). */
int vr_tf_format(const vr_tf_rect* rects, int n, char* out, size_t cap);

/* ---- reference_volume (app/reference_volume.cpp) ---------------------------------------------------- */
/* ctor reference_volume.cpp:11-44: upload (clw_image<short> push) + fetch_stats
 * (opencl_kernels/reference_volume_figures.cl:10-26). */
int vr_volume_upload(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz, vr_volume** out);
/* Asynchronous ingest: returns at once, the copy and fetch_stats run on a copy stream beside the compute stream (upload job
 * k+1 while job k builds its SDF and renders).  `voxels` (pinned memory for a true overlap) must stay valid and unchanged
 * until vr_volume_wait — or any other call that uses the volume, which waits implicitly — returns. */
int vr_volume_upload_async(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz, vr_volume** out);
int vr_volume_wait(vr_volume* vol);
/* the same for voxels that already live in the context's device memory (decoded or generated there): device-to-device copy */
int vr_volume_upload_device(vr_ctx* ctx, const int16_t* device_voxels, int nx, int ny, int nz, vr_volume** out);
void vr_volume_destroy(vr_volume* vol);
/* {min value, max value, min (int)|grad|, max (int)|grad|} over the ORIGINAL volume, unclamped
 * (reference_volume.cpp:33-37).  The clip getters (reference_volume.cpp:82-88) live in the C++ shim. */
int vr_volume_stats(const vr_volume* vol, int32_t out[4]);
/* set_value_clip / set_gradient_clip (reference_volume.cpp:46-52) and get_volume_stats()
 * (reference_volume.cpp:82-88,110-112): out = {max(clip,min_v), min(clip,max_v), max(clip,min_g), min(clip,max_g)} */
int vr_volume_set_value_clip(vr_volume* vol, int lo, int hi);
int vr_volume_set_gradient_clip(vr_volume* vol, int lo, int hi);
int vr_volume_clipped_stats(const vr_volume* vol, float out[4]);
/* dims of the volume the renderer sees (cropped if vr_volume_clip was called) */
int vr_volume_dims(const vr_volume* vol, int out[3]);
/* set_clipping, reference_volume.cpp:54-68 + reference_volume_clip.cl:4-15.  max is exclusive. */
int vr_volume_clip(vr_volume* vol, const uint32_t min[3], const uint32_t max[3]);
/* filter, reference_volume.cpp:70-80 + volume_filter.cl:5-11 + utility_filter.cl:38-62 (5^3 bilateral);
 * replaces the current (cropped or original) volume. */
int vr_volume_filter(vr_volume* vol);
/* current volume back to the host (parity tests) */
int vr_volume_download(const vr_volume* vol, int16_t* out);
/* device pointer of the current volume (x fastest, read-only for the caller; valid until the next clip / filter / destroy) */
const int16_t* vr_volume_device_ptr(const vr_volume* vol);
/* tf_sort_values, histogram.cl:4-32 as launched by renderer.cpp:57-61.
 * range = {min_v, max_v, min_g, max_g}; bins_out = width*height uint32, index x*height + y. */
int vr_histogram(const vr_volume* vol, int width, int height, const float range[4], uint32_t* bins_out);

/* ---- env_map (app/env_map.hpp:10) ------------------------------------------------------------------- */
int vr_envmap_bind(vr_ctx* ctx, const uint8_t* rgba8, int w, int h, vr_envmap** out);
void vr_envmap_destroy(vr_envmap* env);

/* ---- signed_distance_field (app/signed_distance_field.cpp:7-35 + signed_distance_field.cl) ---------- */
/* Scratch during the build: one bit per voxel and level, (max_it - 1) * nx*ny*nz / 8 bytes (2 GiB at 512^3), from the context's
 * stream-ordered pool and returned to it; if that allocation fails the build uses two bit volumes + level planes instead. */
int vr_sdf_build(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out);
void vr_sdf_destroy(vr_sdf* sdf);
/* x-fastest int8, same order as clw_image<char>::pull() (tests/sdf/sdf_test.cpp:24-31) */
int vr_sdf_download(const vr_sdf* sdf, int8_t* out);
/* number of wavefront levels the build ran (diagnostics) */
int vr_sdf_levels(const vr_sdf* sdf);
/* position-weighted 64-bit checksums computed on the device: compare SDF fields / volumes at full size (multi-GPU runs at 1024^3)
 * without moving them to the host */
int vr_sdf_checksum(const vr_sdf* sdf, uint64_t* out);
int vr_volume_checksum(const vr_volume* vol, uint64_t* out);

/* ---- renderer : frame_emitter (app/ui.hpp:29-37, app/renderer.cpp) ---------------------------------- */
/* renderer(ctx), renderer.cpp:8-17 — the reference fixes the frame at SCREEN_WIDTH x SCREEN_HEIGHT
 * (common_defines.hpp:3-4); here the size is a constructor argument. */
int vr_renderer_create(vr_ctx* ctx, int width, int height, vr_renderer** out);
void vr_renderer_destroy(vr_renderer* r);
/* image_set, renderer.cpp:19-23 */
int vr_renderer_set_scene(vr_renderer* r, const vr_volume* vol, const vr_envmap* env);
/* next_event_code_set, renderer.cpp:126-129 — either the clause table or the generated source text */
int vr_renderer_set_tf(vr_renderer* r, const vr_tf_rect* rects, int n_rects);
int vr_renderer_set_tf_code(vr_renderer* r, const char* is_event_gen_src);
/* flush_changes, renderer.cpp:25-43: (re)allocate + zero the voxel cache (buffer_reset.cl:3-13),
 * adopt the pending TF, rebuild the SDF.  Incremental: when only the COLOURS of the transfer function changed since the last
 * flush (same clauses, same value / gradient ranges; same volume, environment map and sampling), the SDF — and the hw-linear
 * step field and textures — are kept, since they depend on the event predicate only; vr_renderer_last_flush_kept_fields tells. */
int vr_renderer_flush(vr_renderer* r);
int vr_renderer_last_flush_kept_fields(const vr_renderer* r);
/* buffer_reset only (renderer.cpp:32-35) */
int vr_renderer_reset_cache(vr_renderer* r);
/* render_frame, renderer.cpp:131-158 + ray_marching.cl:152-199: one sample per pixel accumulated into the
 * voxel cache, then every pixel resolved from the cache.  dir = Position3D(look0, look1, 0, {1,0,0})
 * (renderer.cpp:140), seed = the value std::rand() would have returned (renderer.cpp:142).
 * host_rgba (W*H*4) may be NULL: the frame then stays on the device (no readback, no sync). */
int vr_render_frame(vr_renderer* r, const float pos[3], const float dir[3], int32_t seed, uint8_t* host_rgba);
/* n_frames calls of vr_render_frame with seeds[0..n), reading back only the last frame. */
int vr_render_frames(vr_renderer* r, const float pos[3], const float dir[3], const int32_t* seeds, int n_frames,
                     uint8_t* host_rgba);
/* render_tf, renderer.cpp:45-124 + histogram.cl:34-69; rgba_out = width*height*4.  Histogram ranges are
 * vr_volume_clipped_stats of the bound volume, as in renderer.cpp:57-61. */
int vr_render_tf(vr_renderer* r, int width, int height, uint8_t* rgba_out);
/* renderer-owned pinned host frame (W*H*4): the pointer render_frame() returns in the reference
 * (renderer.cpp:157); passing it as host_rgba makes the readback a single DMA. */
uint8_t* vr_renderer_host_frame(vr_renderer* r);
/* packed voxel cache, 4 ushort per voxel at (nx*nz*y + nx*z + x)*4 (utility.cl:21) — parity tests */
int vr_cache_download(const vr_renderer* r, uint16_t* out);
/* parity checks at sizes where the whole cache is gigabytes: the primary-hit voxel of every pixel as the last trace found it
 * (W*H uint32, 0xFFFFFFFF = environment pixel), and the cache entries of n given voxels (4 ushort each) */
int vr_renderer_hit_download(const vr_renderer* r, uint32_t* out);
int vr_cache_download_at(const vr_renderer* r, const uint32_t* voxels, size_t n, uint16_t* out);
/* the SDF the renderer built at the last flush (renderer.hpp:19) */
const vr_sdf* vr_renderer_sdf(const vr_renderer* r);

/* ---- sampling of the volume and the environment map by the path tracer ----------------------------------------------------
 * Every sampler of the reference requests CLK_FILTER_LINEAR on INTEGER images (utility_ray.cl:130,149, utility_filter.cl:4,
 * utility_environment_map.cl:4), for which OpenCL 1.2 defines no result.
 *   VR_SAMPLING_NEAREST   (default) the filter OpenCL defines for integer images: texel floor(coord), border 0.  All schedules.
 *   VR_SAMPLING_HW_LINEAR what NVIDIA hardware does with the kernels AS SHIPPED (measured through the driver's OpenCL runtime):
 *                         value and gradient taps of get_event_and_value and the environment colour are interpolated by the
 *                         texture unit (texel centres at +0.5, 8-bit weights, border 0 / clamp to edge) and rounded to
 *                         integers; integer-coordinate reads (the SDF read of march) stay texel reads.  All schedules; the
 *                         flush also builds a step field that lets a step skip its seven fetches where the interpolated value
 *                         cannot meet the transfer function (same results; vr_quiet.cu).  The SDF build is unaffected.
 * Takes effect at the next vr_renderer_flush (which copies the volume and the environment map into texture arrays). */
enum { VR_SAMPLING_NEAREST = 0, VR_SAMPLING_HW_LINEAR = 1 };
int vr_renderer_set_sampling(vr_renderer* r, int mode);
/* The same choice for the volume kernels that read through a sampler: fetch_stats (recomputed by this call, so vr_volume_stats /
 * vr_volume_clipped_stats follow), tf_sort_values (vr_histogram, vr_render_tf) and bilateral_filter (vr_volume_filter).  Their value
 * read uses a sampler without an addressing mode (reference_volume_figures.cl:12, histogram.cl:7), which NVIDIA hardware serves like
 * clamp-to-edge.  apply_clip and the SDF build read without a sampler and are the same under both readings. */
int vr_volume_set_sampling(vr_volume* vol, int mode);

/* ---- 2-D frame filter (opencl_kernels/2d_image_filter.cl:6-43 `bilateral_filter(frame, kernel_size, sigma)`) -------
 * The reference ships this kernel but no host code launches it; the call takes the kernel's own arguments.
 *   VR_FILTER2D_REFERENCE : the kernel's arithmetic exactly as written (bit-identical to the source compiled for the CPU):
 *                           gauss() = a/(2*sigma^2) without an exponential, spatial term from pos - off, unsigned colour
 *                           differences, the centre colour accumulated, all channels divided by the red weight sum, alpha 0.
 *                           kernel_size in [0, 64].
 *   VR_FILTER2D_BILATERAL : the bilateral filter it set out to be: exp(-(dx^2+dy^2)/(2 sigma^2)) * exp(-(tap-centre)^2/(2 sigma^2))
 *                           per channel, centre tap included, taps outside the frame skipped, rounded to nearest, alpha kept.
 *                           kernel_size in [0, 15].
 * The reference kernel works in place on a __read_write image (taps race with neighbouring writes); here every tap reads the
 * unfiltered frame.  sigma must be > 0. */
enum { VR_FILTER2D_REFERENCE = 0, VR_FILTER2D_BILATERAL = 1 };
/* filters the renderer's current device frame (the result of the last render_frame / resolve) into a second device buffer
 * (vr_renderer_filtered_device_ptr) and optionally reads that back; the traced frame itself is left untouched */
int vr_renderer_filter_frame(vr_renderer* r, int kernel_size, float sigma, int mode, uint8_t* host_rgba);
void* vr_renderer_filtered_device_ptr(const vr_renderer* r);
/* the same on a caller-supplied RGBA8 image, host to host (rgba_out may equal rgba_in) */
int vr_image_filter(vr_ctx* ctx, const uint8_t* rgba_in, int w, int h, int kernel_size, float sigma, int mode,
                    uint8_t* rgba_out);

/* ---- multi-GPU hooks (no reference counterpart: the reference is single-device) ---------------------- */
/* per-rank token cap for the spp split: 256/R keeps the summed cache under the reference's cap of 256
 * (ray_marching.cl:39) and every 16-bit lane below overflow. Default 256. */
int vr_renderer_set_token_cap(vr_renderer* r, int cap);
/* restrict tracing to the pixel rows [y0,y1) (image-tile split); default whole frame */
int vr_renderer_set_rows(vr_renderer* r, int y0, int y1);
/* device pointers for collectives done by the caller (NCCL / torch.distributed): the packed cache viewed as
 * uint32[2*N] (sum-reducible without cross-lane carries), and the RGBA8 frame. */
void* vr_renderer_cache_device_ptr(const vr_renderer* r);
size_t vr_renderer_cache_bytes(const vr_renderer* r);
void* vr_renderer_frame_device_ptr(const vr_renderer* r);
/* Compact cache exchange for the spp split: all ranks trace the same camera, so they touch the same voxels (the
 * pixels' primary hit voxels).  gather copies cache[hit[pix]] into a W*H*8-byte device buffer (pointer returned by
 * vr_renderer_xchg_device_ptr), the caller sum-reduces that buffer across ranks as int32 words, scatter writes the sums
 * back.  Valid when every rank has traced only this camera since the last vr_renderer_reset_cache / exchange. */
int vr_renderer_xchg_gather(vr_renderer* r);
int vr_renderer_xchg_scatter(vr_renderer* r);
void* vr_renderer_xchg_device_ptr(vr_renderer* r);
size_t vr_renderer_xchg_bytes(const vr_renderer* r);
/* re-run only the resolve pass (after an external cache all-reduce) and optionally read the frame back */
int vr_renderer_resolve(vr_renderer* r, uint8_t* host_rgba);

/* ---- multi-GPU collectives behind the C-ABI (NCCL over NVLink, enqueued on the context's stream) --------
 * One vr_ctx per GPU, one process or thread per vr_ctx.  Rank 0 creates an id (vr_comm_unique_id) and hands it to the other
 * ranks by whatever means the host has (a file, a socket, MPI, a torch store); every rank then calls vr_comm_init.  Without a
 * communicator (or with one rank) every call below degenerates to its single-GPU meaning. */
#define VR_COMM_ID_BYTES 128
int vr_comm_unique_id(uint8_t id[VR_COMM_ID_BYTES]);
int vr_comm_init(vr_ctx* ctx, int rank, int nranks, const uint8_t id[VR_COMM_ID_BYTES]);
void vr_comm_destroy(vr_ctx* ctx); /* also done by vr_ctx_destroy */
int vr_comm_rank(const vr_ctx* ctx);
int vr_comm_size(const vr_ctx* ctx);
int vr_comm_barrier(vr_ctx* ctx);
/* small host-side reductions for drivers (timings, checksums): dtype 0 int32 / 1 uint32 / 2 float64, op 0 sum / 1 min / 2 max */
int vr_comm_allreduce_host(vr_ctx* ctx, void* values, int count, int dtype, int op);
/* the z-partition every sharded call uses: planes [z0, z1) of rank `rank`, multiples of 8 planes (brick layers of the SDF) */
int vr_comm_slab(const vr_ctx* ctx, int nz, int rank, int* z0, int* z1);
/* spp split (BASELINE config 3): every rank has traced its own seeds of the SAME camera into its own cache with a token cap of
 * 256/N (vr_renderer_set_token_cap).  Sums the touched cache entries over the ranks — one 8-byte entry per shaded pixel,
 * numbered in pixel order, as uint32 words (no carries between the 16-bit lanes by construction) — writes the sums back into
 * every rank's cache and resolves the frame from them (ray_marching.cl:82-99); host_rgba may be NULL.  When the ranks' caps add
 * up to more than 256 (N ranks that each accumulate a full 64-spp job: weak scaling) the lanes travel as four uint32 words, the
 * frame is resolved from those wide sums and the per-rank caches keep their partial sums. */
int vr_cache_allreduce(vr_renderer* r, uint8_t* host_rgba);
/* image-tile split (BASELINE config 4): rows are dealt out in blocks of block_rows, block b is traced by rank b % nranks;
 * every rank keeps its own cache (visible voxels of its rows).  vr_frame_allgather completes the frame on every rank. */
int vr_renderer_set_row_blocks(vr_renderer* r, int block_rows, int rank, int nranks);
int vr_frame_allgather(vr_renderer* r, uint8_t* host_rgba);
/* sharded ingest: rank r passes only ITS planes [z0, z1) of vr_comm_slab (nx*ny*(z1-z0) voxels, preferably pinned); the other
 * planes arrive from the other ranks over NVLink instead of N copies of the whole volume over PCIe.  Result: the whole volume
 * on every rank, as after vr_volume_upload. */
int vr_volume_upload_sharded(vr_ctx* ctx, const int16_t* own_planes, int nx, int ny, int nz, vr_volume** out);
/* the same without blocking (cf. vr_volume_upload_async): copy, gather (on a second communicator, so that it can run beside the
 * compute stream's collectives) and fetch_stats go to the copy stream; every call that uses the volume waits for it */
int vr_volume_upload_sharded_async(vr_ctx* ctx, const int16_t* own_planes, int nx, int ny, int nz, vr_volume** out);
/* z-slab SDF build (BASELINE config 5): every rank runs the wavefront on its slab + 16 halo planes, swaps the boundary planes
 * of the bit volume with its z-neighbours every 14 levels and gathers the field; bit-identical to vr_sdf_build.  `vol` is the
 * whole volume (replicated).  vr_renderer_set_sharded_build lets vr_renderer_flush build its SDF this way where slabs pay: a
 * volume whose levels are a single wave of thread blocks (up to 512^3 on a B200) costs the same per level however thin the slab,
 * so it is built on every rank without any collective; larger volumes are built in slabs (1024^3: 13.6 -> 5.7 ms on 8 GPUs). */
int vr_sdf_build_sharded(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out);
int vr_renderer_set_sharded_build(vr_renderer* r, int enable);
/* the same build without the final gather: only the planes vr_comm_slab assigns to this rank are valid in the result (a consumer
 * that works slab by slab, or writes its slab out, does not need the other ranks' planes; vr_sdf_download returns the whole
 * array, meaningful in those planes only).  Not for rendering. */
int vr_sdf_build_slab_only(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out);
/* tf_sort_values / bilateral_filter over the rank's planes of the replicated volume + all-reduce of the bins / gather of the
 * filtered planes: same results as vr_histogram / vr_volume_filter on every rank */
int vr_histogram_sharded(const vr_volume* vol, int width, int height, const float range[4], uint32_t* bins_out);
int vr_volume_filter_sharded(vr_volume* vol);

/* ---- z-slab sharding (multi-GPU; no reference counterpart — SURVEY.md 8e).  A rank holds planes [z0-h, z1+h) of the volume:
 * vr_volume_upload_slab marks [z_lo, z_hi) (indices into the slab) as its own; stats and histogram count only those, the
 * other planes serve gradient / filter taps and the SDF wave.  Partial results combine with MIN/MAX (stats) and SUM (bins). */
int vr_volume_upload_slab(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz_ext, int z_lo, int z_hi, vr_volume** out);
int vr_volume_download_planes(const vr_volume* vol, int z0, int nplanes, int16_t* out);
/* SDF of a slab, level by level.  The wave's results go stale from the slab's artificial ends inwards by one plane per
 * level (plus two at the start), so the caller runs K levels (vr_sdf_slab_advance), overwrites the halo planes of the
 * CURRENT bit volume (vr_sdf_slab_bits: uint32 [nz_ext][ny][ceil(nx/32)], one bit per voxel already reached) with the
 * neighbours' interior planes, calls vr_sdf_slab_mark_imported and continues until vr_sdf_slab_finished.
 * max_it_global = min(max(global dims)/2, 127) (signed_distance_field.cpp:11). */
typedef struct vr_sdf_slab vr_sdf_slab;
int vr_sdf_slab_create(vr_ctx* ctx, const vr_volume* ext_slab, const vr_tf_rect* rects, int n_rects, int max_it_global,
                       vr_sdf_slab** out);
int vr_sdf_slab_advance(vr_sdf_slab* s, int nlevels, int* levels_done);
void* vr_sdf_slab_bits(vr_sdf_slab* s);
size_t vr_sdf_slab_plane_words(const vr_sdf_slab* s);
int vr_sdf_slab_mark_imported(vr_sdf_slab* s);
int vr_sdf_slab_finished(const vr_sdf_slab* s);
int vr_sdf_slab_download(vr_sdf_slab* s, vr_ctx* ctx, int nx, int ny, int nz_ext, int z0, int nplanes, int8_t* out);
void vr_sdf_slab_destroy(vr_sdf_slab* s);

/* ---- instrumentation ---------------------------------------------------------------------------------- */
/* When enabled the trace kernel also accumulates the per-sample counters of SURVEY.md §8(d):
 * {march steps, shading normals, env fetches, primary hits, admitted samples, samples}. */
int vr_renderer_enable_counters(vr_renderer* r, int enable);
int vr_renderer_counters(const vr_renderer* r, uint64_t out[6], int reset);
/* Scheduling of the trace phase — same per-sample computation and results in all modes:
 *   2 (default) primary reuse: everything ray_marching.cl does before it first uses random_seed (ray generation, box cut,
 *               primary march, environment colour of escaping rays, hit voxel, shading normal — :162-170, :21-33, :42)
 *               depends on the camera only and is evaluated once per pixel and call (k_primary); token admission and the
 *               secondary paths run per (pixel, frame) on persistent warps (k_trace_pt)
 *   1 hybrid:   a dense thread-per-pixel primary phase per frame (k_trace) queues the admitted hits, persistent warps
 *               (k_trace_pt) run their secondary paths in refilled lanes
 *   0 one thread per pixel and frame for its whole life (k_trace alone) */
int vr_renderer_set_trace_mode(vr_renderer* r, int mode);
/* Primary reuse across calls (mode 2 only): 1 (default) = within one vr_render_frames call only; 2 = also across calls
 * while camera, rows and scene are unchanged — for the reference's usage, one render_frame call per sample
 * (renderer.cpp:131-158).  Flush, scene, row-range or camera changes always re-march. */
int vr_renderer_set_primary_reuse(vr_renderer* r, int level);
/* When enabled, every trace / resolve launch is bracketed by CUDA events on the context's stream;
 * vr_renderer_kernel_times synchronises and returns the summed device time of the trace kernel (out_ms[0]) and of
 * the resolve kernel (out_ms[1]) and the number of frames measured since the last reset. */
int vr_renderer_enable_timing(vr_renderer* r, int enable);
int vr_renderer_kernel_times(vr_renderer* r, double out_ms[2], int* n_frames, int reset);
/* Schedule tuning of the persistent-warp tracer; never changes a result.  Keys: "pixel_major" (items per pixel group, 0 =
 * frame-major), "rule_a"/"rule_b" (leave the march region when marching*a < waiting*b), "steps_per_check", "lin_sched" (hw-linear
 * march region: 0 two loops with leave rules, 1 weighted choice per round) with "lin_w_fast" / "lin_w_slow" / "lin_w_event", "pt_ctas" (register budget;
 * only in the A/B build, tools/ab). */
int vr_renderer_set_tuning(vr_renderer* r, const char* key, int value);
/* hw-linear step field: the quiet-octant byte of every voxel cell, x fastest (tests: equals the oracle's orc_quiet_cells) */
int vr_renderer_quiet_download(const vr_renderer* r, uint8_t* out);
/* hw-linear fetch known answers: the value get_event_and_value reads at n float positions (x, y, z triples) through the
 * renderer's volume texture (tests: equals the oracle's model of the hardware filter, itself pinned on 874 545 OpenCL samples) */
int vr_debug_linear_fetch(const vr_renderer* r, const float* xyz, int n, int32_t* out);
/* Device RNG known answers (utility_sampling.cl:13-21,40-50): runs the device functions of the trace kernels on n items
 * (seed, gid0, gid1, normal.xyz + roughness) and returns the three hashes `ra`, the components `(ra % 2048) - 1024`, and the
 * sampled direction. */
int vr_debug_rng_dump(vr_ctx* ctx, const int32_t* seeds, const uint32_t* gid_xy, const float* normal_rough, int n, int32_t* ra_out,
                      int32_t* comp_out, float* dir_out);

#ifdef __cplusplus
}
#endif
#endif /* VR_H */
