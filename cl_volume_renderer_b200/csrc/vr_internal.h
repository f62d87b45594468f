// vr_internal.h — host-side object model behind the opaque handles of include/vr.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/vr.h"

void vr_set_error(const char* fmt, ...);

#define VR_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      vr_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VR_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define VR_REQUIRE(cond, msg)                                     \
  do {                                                            \
    if (!(cond)) {                                                \
      vr_set_error("%s:%d: %s", __FILE__, __LINE__, msg);         \
      return VR_ERR_INVALID;                                      \
    }                                                             \
  } while (0)

#define VR_TRY(call)          \
  do {                        \
    int s__ = (call);         \
    if (s__ != VR_OK) return s__; \
  } while (0)

struct vr_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // vr_volume_upload_async: host->device copy + fetch_stats beside the compute stream
  cudaStream_t aux_stream = nullptr;   // vr_renderer_flush: the (bandwidth-bound) cache reset beside the (latency-bound) SDF build
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint64_t launches = 0;
  int32_t* scratch = nullptr;  // small device scratch (counters, stats), 4 KiB
  int32_t* scratch_host = nullptr;  // pinned mirror
  // pinned host frame buffers are recycled across renderers: cudaMallocHost / cudaFreeHost cost milliseconds each
  struct PinnedBuf { void* p; size_t bytes; bool in_use; };
  std::vector<PinnedBuf> pinned;
  // 3-D arrays (+ surface objects) of the SDF (8 bits per voxel) and of the hw-linear step field (16 bits) are recycled by size
  // as well: cudaMalloc3DArray / cudaFreeArray synchronise.  At most four unused arrays are kept (array3d_release).
  struct Array3D { cudaArray_t arr; cudaSurfaceObject_t surf; int nx, ny, nz, bits; bool in_use; uint64_t released; };
  std::vector<Array3D> arrays3d;
  uint64_t array_clock = 0;
  void* comm = nullptr;  // ncclComm_t once vr_comm_init was called (vr_comm.cu)
  void* comm_copy = nullptr;  // a second communicator over the same ranks for collectives on the copy stream (sharded async ingest)
  int comm_rank = 0, comm_size = 1;
};
int array3d_acquire(vr_ctx* ctx, int nx, int ny, int nz, int bits, cudaArray_t* arr, cudaSurfaceObject_t* surf);
void array3d_release(vr_ctx* ctx, cudaArray_t arr);

// Device-side TF table, passed to kernels by value.
struct TfTable {
  vr_tf_rect r[VR_TF_MAX_RECTS];
  float e[VR_TF_MAX_RECTS][4];  // rgba / 255.0f (IEEE, evaluated on the host)
  int n;
  int needs_gradient;  // any clause carries a gradient test
};

struct vr_volume {
  vr_ctx* ctx = nullptr;
  int16_t* original = nullptr;  // device, x fastest
  int onx = 0, ony = 0, onz = 0;
  int16_t* cropped = nullptr;  // device, non-null once clipped (reference_volume.hpp:31-33)
  int nx = 0, ny = 0, nz = 0;  // dims of the current volume
  int32_t stats[4] = {0, 0, 0, 0};
  int32_t raw_range[2] = {0, -1};  // min / max of the voxel values as last measured under NEAREST (bounds the box-averaged copy); max < min: unknown
  // vr_volume_upload_async: copy + stats are in flight on ctx->copy_stream until `ready`; every API call that uses the
  // volume finishes the upload first (volume_finish in vr_api.cu)
  cudaEvent_t ready = nullptr;
  bool pending = false;
  int32_t* stats_dev = nullptr;
  int32_t* stats_pin = nullptr;
  int zlo = 0, zhi = 0;  // planes the stats / histogram cover (whole volume unless uploaded as a z-slab with halo planes)
  // vr_volume_set_sampling(VR_SAMPLING_HW_LINEAR): the box-averaged volume (vr_volume_ops.cu k_boxavg: what the hardware's linear
  // filter returns at integer coordinates), box_px x (ny+1) x (nz+1), rebuilt whenever the current volume changes (clip, filter)
  int sampling = VR_SAMPLING_NEAREST;
  int16_t* box = nullptr;
  int box_px = 0;
  // bumped whenever the current volume changes (clip, filter): a renderer whose SDF / cache / textures were built from an older
  // generation refuses to trace until it is flushed again (the reference always flushes after set_clipping, ui.cpp:273-278)
  uint64_t generation = 1;
  int value_clip[2] = {INT32_MIN, INT32_MAX};     // reference_volume.hpp:35-36
  int gradient_clip[2] = {INT32_MIN, INT32_MAX};
  const int16_t* current() const { return cropped ? cropped : original; }
  size_t count() const { return (size_t)nx * ny * nz; }
};

struct vr_envmap {
  vr_ctx* ctx = nullptr;
  uchar4* texels = nullptr;
  int w = 0, h = 0;
};

struct vr_sdf {
  vr_ctx* ctx = nullptr;
  int8_t* field = nullptr;  // device, 8x8x8 bricks of 512 contiguous bytes, dims padded up to multiples of 8
  int nx = 0, ny = 0, nz = 0;
  int levels = 0;
  int max_it = 0;
  // the same int8 values in a 3-D CUDA array behind a surface object: hardware addressing and zero border for the per-step
  // gather of k_trace_pt (VR_SDF_SURF=0 turns it off; the bricked field stays the source for everything else)
  cudaArray_t arr = nullptr;
  cudaSurfaceObject_t surf = 0;
};

struct vr_renderer {
  vr_ctx* ctx = nullptr;
  int W = 0, H = 0;
  int row0 = 0, row1 = 0;
  const vr_volume* vol = nullptr;
  const vr_envmap* env = nullptr;
  TfTable tf_pending{};
  TfTable tf_active{};
  bool have_tf = false;
  vr_sdf* sdf = nullptr;
  uint32_t* cache = nullptr;  // 2 x uint32 per voxel: lo = R | G<<16, hi = B | tokens<<16 (utility.cl:39-54)
  size_t cache_voxels = 0;
  uint32_t* hit = nullptr;    // per pixel: voxel number of the primary hit, 0xFFFFFFFF = env pixel
  uchar4* frame = nullptr;    // device RGBA8
  uint8_t* frame_host = nullptr;  // pinned staging (used when the caller's buffer is pageable)
  int token_cap = 256;
  // vr_renderer_set_sampling: VR_SAMPLING_HW_LINEAR traces through texture objects built by the flush (copies of the current
  // volume and of the environment map in CUDA arrays)
  int sampling = VR_SAMPLING_NEAREST;
  cudaArray_t vol_arr = nullptr, env_arr = nullptr;
  cudaTextureObject_t vol_tex = 0, env_tex = 0;
  int tex_dims[5] = {0, 0, 0, 0, 0};  // volume and env-map sizes the arrays were allocated for
  cudaArray_t lin_arr = nullptr;      // the step field of the hw-linear path (vr_quiet.cu), 16 bits per voxel cell
  cudaSurfaceObject_t lin_surf = 0;
  // what the last flush was built from: tracing with anything else bound is refused (vr_renderer_check_flushed)
  const vr_volume* flushed_vol = nullptr;
  const vr_envmap* flushed_env = nullptr;
  uint64_t flushed_generation = 0;
  int flushed_sampling = -1;
  bool flushed_sharded = false;
  bool fields_kept = false;           // the last flush kept SDF and step field (same event predicate: only colours changed)
  uchar4* filtered = nullptr;         // vr_renderer_filter_frame writes here: the traced frame stays as it is
  // schedule tuning (vr_renderer_set_tuning): never changes a result
  struct Tuning {
    int pixel_major = 1;               // k_trace_pt<.., REUSE> item order: groups of this many pixels x all frames (0: frame-major)
    int rule[2] = {5, 1};              // k_trace_pt leaves its march region when marching lanes * rule[0] < waiting lanes * rule[1]
    int lin_sched = 0;                 // hw-linear: 0 quiet-step loop + event-test loop with leave rules, 1 weighted choice per round
    int lin_w[3] = {3, 2, 2};          // their parameters (quiet steps / event tests / event processing)
    int spc = 4;                       // steps per scheduling decision of k_trace_pt
    int sm_k = 16, sm_leave = 16;      // k_trace_sm (trace mode 3): steps per visit of a marching batch, early-stop threshold
    int surf = 1;                      // NEAREST k_trace_pt gathers the SDF through the surface object (0: bricked field, __ldg)
    int pt_ctas = 0;                   // -DVR_AB builds only: register budget variant of k_trace_pt
    int pt_slots = 1;                  // sample slots per lane of the NEAREST production trace: 1 = k_trace_pt, 2 = k_trace_pt2
    int pt2_ctas = 8;                  // CTAs per SM (register budget) of k_trace_pt2: 6, 7, 8, 10
  } tune;
  // Which cache entries can be non-zero: 0 none (just reset), 1 only cache[hit[pix]] of the current `hit` buffer (every trace
  // since the last reset used the camera / rows in dirty_pos.. below), 2 unknown (full reset needed).  A frame reset then
  // clears W*H entries instead of 8 bytes x voxels (vr_renderer_reset_cache).
  int cache_dirty = 2;
  bool cache_exposed = false;  // vr_renderer_cache_device_ptr handed the raw pointer out: writes can no longer be tracked
  float dirty_pos[3] = {0, 0, 0}, dirty_dir[3] = {0, 0, 0};
  int dirty_rows[2] = {0, 0};
  bool count = false;
  unsigned long long* counters = nullptr;  // 6 x u64 on device (+ 2 spare words: [6] is k_trace_pt's work counter)
  // 0: k_trace alone (a thread per pixel for its whole life), 1: hybrid k_trace + k_trace_pt per frame,
  // 2 (default): k_primary once per pixel and call + k_trace_pt per (pixel, frame); 3: the same with k_trace_sm (slots in
  // shared memory, packed batches)
  int trace_mode = 2;
  bool primary_valid = false;         // r->queue holds the primary records of (primary_pos, primary_dir, primary_rows)
  bool primary_across_calls = false;  // vr_renderer_set_primary_reuse(r, 2)
  float primary_pos[3] = {0, 0, 0}, primary_dir[3] = {0, 0, 0};
  int primary_rows[2] = {0, 0};
  // Incremental pull of the renderer-owned host frame (vr_renderer_host_frame): while the primary records stay valid the
  // environment pixels of the frame cannot change, so a pull copies only the bounding box of the shaded pixels (k_primary
  // reduces it, bbox_pin receives it) instead of the whole frame.  primary_epoch counts k_primary runs; host_epoch is the epoch
  // whose complete frame the host buffer holds (0: none).
  uint64_t primary_epoch = 0, host_epoch = 0;
  int* bbox_dev = nullptr;   // {min x, min y, max x, max y} of the shaded pixels
  int* bbox_pin = nullptr;
  uint4* queue = nullptr;  // hybrid schedule: admitted primary hits (3 x uint4 each)
  size_t queue_cap = 0;
  uint2* xchg = nullptr;  // W*H compact cache entries for the spp-split exchange (allocated on first use)
  // vr_cache_allreduce needs the number of shaded pixels on the host (the NCCL count).  It is a function of the hit buffer,
  // i.e. of camera, rows and flushed scene: remembered per such signature, so a progressive loop reads it back once
  bool xc_sig_valid = false;
  unsigned xc_count_host = 0;
  float xc_pos[3] = {0, 0, 0}, xc_dir[3] = {0, 0, 0};
  uint64_t xc_flush = 0, flush_count = 0;
  uint32_t* xc_idx = nullptr;    // per shaded pixel: its entry in the dense exchange array (vr_cache_allreduce)
  unsigned* xc_counts = nullptr; // per 256-pixel block: shaded pixels before it; [blocks] = total
  // image-tile split: rows are dealt out in blocks of blk_rows, block b belongs to rank b % blk_n (vr_renderer_set_row_blocks)
  int blk_rows = 0, blk_rank = 0, blk_n = 1;
  uint32_t* gather_buf = nullptr;  // vr_frame_allgather staging: (1 + ranks) chunks
  size_t gather_bytes = 0;
  bool sharded_build = false;      // vr_renderer_set_sharded_build: the flush builds the SDF z-slab-sharded over the communicator
  bool timing = false;
  std::vector<cudaEvent_t> ev;  // 3 events per timed launch: before trace, between, after resolve
  size_t ev_used = 0;
  std::vector<int> ev_frames;   // frames traced by each timed launch
};

// ---- kernel launchers (defined in the .cu files) -----------------------------------------------------
// stats / histogram over the planes [zlo, zhi) only (z-slab sharding: the other planes are halo for the gradient taps)
int vrk_fetch_stats(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int32_t out[4], int zlo, int zhi);
// the same on any stream, without waiting: dev4 / pin4 = 4 ints of device scratch / pinned host memory that receive the result
int vrk_fetch_stats_enqueue(vr_ctx* ctx, cudaStream_t stream, const int16_t* vol, int nx, int ny, int nz, int zlo, int zhi,
                            int32_t* dev4, int32_t* pin4);
// completes an enqueued fetch_stats once its transfer into pin4 has finished (accepts the integer kernel's result or reruns in fp32)
int vrk_fetch_stats_finalize(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int zlo, int zhi, const int32_t* pin4, int32_t out[4]);
int vrk_clip(vr_ctx* ctx, const int16_t* src, int snx, int sny, int snz, const uint32_t start[3], int16_t* dst, int nx,
             int ny, int nz);
int vrk_bilateral(vr_ctx* ctx, const int16_t* src, int16_t* dst, int nx, int ny, int nz);
int vrk_histogram(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int width, int height, const float range[4],
                  uint32_t* bins_dev, int zlo, int zhi, int vol_min_value = 0, int vol_max_value = -0x7fffffff);  // max < min: range unknown
struct vr_sdf_slab;
int vrk_sdf_slab_create(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int max_it, vr_sdf_slab** out);
int vrk_sdf_slab_advance(vr_sdf_slab* s, int nlevels, int* done);
uint32_t* vrk_sdf_slab_bits(vr_sdf_slab* s);
size_t vrk_sdf_slab_plane_words(const vr_sdf_slab* s);
void vrk_sdf_slab_mark_imported(vr_sdf_slab* s);
int vrk_sdf_slab_level(const vr_sdf_slab* s);
bool vrk_sdf_slab_finished(const vr_sdf_slab* s);
int vrk_sdf_slab_assemble(vr_sdf_slab* s, int8_t* field, cudaSurfaceObject_t surf = 0);
int vrk_sdf_slab_status(vr_sdf_slab* s);
bool vrk_sdf_single_wave(const vr_ctx* ctx, int nx, int ny, int nz);
void vrk_sdf_slab_destroy(vr_sdf_slab* s);
int vrk_tf_image(vr_ctx* ctx, int32_t* bins_dev, int* scratch_dev, int width, int height, uchar4* out_dev);
// surf != 0: the field is also written into that surface (the production schedule does it in its assembly pass)
int vrk_sdf_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field,
                  int* levels_out, int* max_it_out, cudaSurfaceObject_t surf = 0);
// the same field built z-slab-sharded over the context's communicator (vr_comm.cu); falls back to vrk_sdf_build without one
int vrk_sdf_build_sharded(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field, int* levels_out,
                          int* max_it_out, cudaSurfaceObject_t surf, bool gather);
size_t vrk_sdf_field_bytes(int nx, int ny, int nz);
int vrk_sdf_unbrick(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, int8_t* linear);
int vrk_sdf_to_surface(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, cudaSurfaceObject_t surf);
int vrk_cache_reset(vr_ctx* ctx, uint32_t* cache, size_t voxels, cudaStream_t stream = nullptr);  // nullptr: the context's stream
int vrk_cache_reset_hits(vr_ctx* ctx, uint32_t* cache, const uint32_t* hit, size_t pixels);
#define VR_MAX_BATCH 64
// trace `nframes` frames (seeds[0..nframes)) in ONE launch (gridDim.z = frame), then optionally resolve once
int vrk_render(vr_renderer* r, const float pos[3], const float dir[3], const int32_t* seeds, int nframes, bool trace,
               bool resolve, bool first_of_call = true);

int vrk_xchg(vr_renderer* r, uint2* xchg, bool scatter);
int vrk_xc_gather(vr_renderer* r, unsigned** count_dev, bool wide);
int vrk_xc_scatter_resolve(vr_renderer* r, bool wide);
int vrk_checksum(vr_ctx* ctx, const void* dev, size_t bytes, uint64_t* out);
int vrk_cache_gather(vr_ctx* ctx, const uint32_t* cache, const uint32_t* idx_dev, size_t n, uint2* out_dev);
void vr_comm_release(vr_ctx* ctx);
// sharded: 0 single-GPU build, 1 z-slab build + gather of the field, 2 z-slab build only (the rank's own planes are valid)
int sdf_build_impl(vr_ctx* ctx, const vr_volume* vol, const TfTable& tf, vr_sdf** out, int sharded);
int volume_finish(const vr_volume* cv);
cudaError_t pinned_acquire(vr_ctx* ctx, void** p, size_t bytes);
void pinned_release(vr_ctx* ctx, void* p);
// hw-linear step field: per voxel cell the SDF byte + 8 quiet-octant bits, written into a 16-bit 3-D surface (vr_quiet.cu)
int vrk_lin_field_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const int8_t* sdf_bricked, const TfTable& tf,
                        cudaSurfaceObject_t out);
int vrk_lin_field_masks(vr_ctx* ctx, cudaSurfaceObject_t field, int nx, int ny, int nz, uint8_t* masks_dev);
// device RNG known-answer dump (tests): hemisphere integer triples and directions over a (seed, gid) grid (vr_render.cu)
int vrk_linear_fetch(vr_ctx* ctx, cudaTextureObject_t tex, const float* xyz_dev, int n, int32_t* out_dev);
int vrk_rng_dump(vr_ctx* ctx, const int32_t* seeds_dev, const uint32_t* gid_dev, int n, const float* normal_rough_dev, int32_t* ra_dev,
                 int32_t* comp_dev, float* dir_dev);
// the volume kernels under VR_SAMPLING_HW_LINEAR: box = the box-averaged volume, px its padded row length (vr_volume_ops.cu)
int vrk_boxavg(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int16_t* out, int px);
int vrk_fetch_stats_linear(vr_ctx* ctx, const int16_t* box, int px, const int16_t* vol, int nx, int ny, int nz, int32_t out[4], int zlo,
                           int zhi);
int vrk_histogram_linear(vr_ctx* ctx, const int16_t* box, int px, const int16_t* vol, int nx, int ny, int nz, int width, int height,
                         const float range[4], uint32_t* bins_dev, int zlo, int zhi, int vol_min_value, int raw_min = 0, int raw_max = -1);
int vrk_bilateral_linear(vr_ctx* ctx, const int16_t* box, int px, int nx, int ny, int nz, int16_t* dst);
// 2d_image_filter.cl: src != dst, both w*h RGBA8 on the device
int vrk_filter2d(vr_ctx* ctx, const uchar4* src, uchar4* dst, int w, int h, int kernel_size, float sigma, int mode);
TfTable vr_make_tf_table(const vr_tf_rect* rects, int n);
