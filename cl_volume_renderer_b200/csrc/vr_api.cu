// vr_api.cu — the extern "C" surface of include/vr.h: handle lifetime, uploads/downloads and the host-side
// sequencing of the reference's renderer / reference_volume / signed_distance_field classes.
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <new>
#include "vr_internal.h"

static thread_local char g_err[1024] = "";

void vr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* vr_last_error(void) { return g_err; }

// Device memory comes from the device's default stream-ordered pool, whose release threshold is raised to "never" in
// vr_ctx_create: scene objects that are destroyed and re-created (volume upload, TF flush, SDF rebuild) then recycle
// their blocks without a driver round trip.  Frees are ordered on the context's stream.
template <typename T>
static cudaError_t pool_alloc(vr_ctx* ctx, T** p, size_t bytes) {
  return cudaMallocAsync(reinterpret_cast<void**>(p), bytes ? bytes : 1, ctx->stream);
}
static void pool_free(vr_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}

cudaError_t pinned_acquire(vr_ctx* ctx, void** p, size_t bytes) {
  for (auto& b : ctx->pinned)
    if (!b.in_use && b.bytes == bytes) { b.in_use = true; *p = b.p; return cudaSuccess; }
  cudaError_t e = cudaMallocHost(p, bytes);
  if (e == cudaSuccess) ctx->pinned.push_back({*p, bytes, true});
  return e;
}
void pinned_release(vr_ctx* ctx, void* p) {
  for (auto& b : ctx->pinned)
    if (b.p == p) b.in_use = false;
}

// 3-D CUDA arrays (SDF surface, hw-linear step field and volume texture), recycled by (size, kind).  Releasing keeps at most four
// unused arrays (the flush builds the fresh SDF before it lets go of the old one; a hw-linear renderer holds three kinds);
// older ones are freed, so a session that walks through many clip boxes does not accumulate an array per size.
int array3d_acquire(vr_ctx* ctx, int nx, int ny, int nz, int bits, cudaArray_t* arr, cudaSurfaceObject_t* surf) {
  for (auto& a : ctx->arrays3d)
    if (!a.in_use && a.nx == nx && a.ny == ny && a.nz == nz && a.bits == bits) { a.in_use = true; *arr = a.arr; *surf = a.surf; return VR_OK; }
  // bits: 8 = int8 + surface (SDF), 16 = uint16 + surface (hw-linear step field), 17 = int16, texture only (hw-linear volume copy)
  cudaChannelFormatDesc desc = bits == 8 ? cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindSigned)
                               : (bits == 16 ? cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned)
                                             : cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindSigned));
  cudaArray_t a = nullptr;
  cudaSurfaceObject_t so = 0;
  cudaError_t e = cudaMalloc3DArray(&a, &desc, make_cudaExtent(nx, ny, nz), bits == 17 ? cudaArrayDefault : cudaArraySurfaceLoadStore);
  if (e == cudaSuccess && bits != 17) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = a;
    e = cudaCreateSurfaceObject(&so, &rd);
  }
  if (e != cudaSuccess) {
    if (a) cudaFreeArray(a);
    vr_set_error("3-D array %dx%dx%d (%d bits): %s", nx, ny, nz, bits, cudaGetErrorString(e));
    return VR_ERR_CUDA;
  }
  ctx->arrays3d.push_back({a, so, nx, ny, nz, bits, true, 0});
  *arr = a; *surf = so;
  return VR_OK;
}
void array3d_release(vr_ctx* ctx, cudaArray_t arr) {
  if (!arr) return;
  for (auto& a : ctx->arrays3d)
    if (a.arr == arr) { a.in_use = false; a.released = ++ctx->array_clock; }
  for (;;) {
    int unused = 0, oldest = -1;
    for (size_t i = 0; i < ctx->arrays3d.size(); ++i)
      if (!ctx->arrays3d[i].in_use) {
        ++unused;
        if (oldest < 0 || ctx->arrays3d[i].released < ctx->arrays3d[oldest].released) oldest = (int)i;
      }
    if (unused <= 4) break;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->arrays3d[oldest].surf) cudaDestroySurfaceObject(ctx->arrays3d[oldest].surf);
    cudaFreeArray(ctx->arrays3d[oldest].arr);
    ctx->arrays3d.erase(ctx->arrays3d.begin() + oldest);
  }
}

// ---- context -------------------------------------------------------------------------------------------------
extern "C" int vr_ctx_create(int device_ordinal, vr_ctx** out) {
  VR_REQUIRE(out, "vr_ctx_create: null out");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    vr_set_error("vr_ctx_create: no usable CUDA device (%s) — this library has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return VR_ERR_CUDA;
  }
  VR_REQUIRE(device_ordinal >= 0 && device_ordinal < count, "vr_ctx_create: device ordinal out of range");
  VR_CUDA(cudaSetDevice(device_ordinal));
  cudaDeviceProp prop;
  VR_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  if (prop.major != 10) {
    vr_set_error("vr_ctx_create: device %d is sm_%d%d; this build contains sm_100a code only", device_ordinal,
                 prop.major, prop.minor);
    return VR_ERR_CUDA;
  }
  vr_ctx* c = new (std::nothrow) vr_ctx();
  if (!c) return VR_ERR_NOMEM;
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  cudaMemPool_t pool;
  const uint64_t keep = UINT64_MAX;
  e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaDeviceGetDefaultMemPool(&pool, device_ordinal);
  if (e == cudaSuccess) e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, const_cast<uint64_t*>(&keep));
  if (e == cudaSuccess) e = cudaMalloc(&c->scratch, 4096);
  if (e == cudaSuccess) e = cudaMallocHost(&c->scratch_host, 4096);
  if (e != cudaSuccess) {
    vr_set_error("vr_ctx_create: %s", cudaGetErrorString(e));
    vr_ctx_destroy(c);  // releases whatever was created
    return VR_ERR_CUDA;
  }
  *out = c;
  return VR_OK;
}

extern "C" void vr_ctx_destroy(vr_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->scratch) cudaFree(c->scratch);
  if (c->scratch_host) cudaFreeHost(c->scratch_host);
  for (auto& b : c->pinned) cudaFreeHost(b.p);
  for (auto& a : c->arrays3d) { if (a.surf) cudaDestroySurfaceObject(a.surf); cudaFreeArray(a.arr); }
  vr_comm_release(c);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" int vr_ctx_synchronize(vr_ctx* c) {
  VR_REQUIRE(c, "vr_ctx_synchronize: null ctx");
  VR_CUDA(cudaStreamSynchronize(c->stream));
  return VR_OK;
}

extern "C" void* vr_ctx_stream(vr_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" uint64_t vr_ctx_launch_count(const vr_ctx* c) { return c ? c->launches : 0; }
extern "C" int vr_ctx_array_count(const vr_ctx* c) { return c ? (int)c->arrays3d.size() : 0; }

// ---- volume --------------------------------------------------------------------------------------------------
// completes a vr_volume_upload_async: waits for the copy stream's event, takes the stats, releases the staging objects
int volume_finish(const vr_volume* cv) {
  vr_volume* v = const_cast<vr_volume*>(cv);
  if (!v || !v->pending) return VR_OK;
  VR_CUDA(cudaEventSynchronize(v->ready));
  v->pending = false;
  VR_TRY(vrk_fetch_stats_finalize(v->ctx, v->original, v->nx, v->ny, v->nz, v->zlo, v->zhi, v->stats_pin, v->stats));
  pinned_release(v->ctx, v->stats_pin);
  v->stats_pin = nullptr;
  pool_free(v->ctx, v->stats_dev);
  v->stats_dev = nullptr;
  cudaEventDestroy(v->ready);
  v->ready = nullptr;
  return VR_OK;
}

static int volume_upload_impl(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz, int zlo, int zhi, vr_volume** out) {
  VR_REQUIRE(ctx && voxels && out, "vr_volume_upload: null argument");
  VR_REQUIRE(nx > 0 && ny > 0 && nz > 0, "vr_volume_upload: dimensions must be positive");
  VR_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 32) - 1, "vr_volume_upload: more than 2^32-2 voxels");
  VR_REQUIRE(zlo >= 0 && zlo < zhi && zhi <= nz, "vr_volume_upload_slab: bad interior plane range");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_volume* v = new (std::nothrow) vr_volume();
  if (!v) return VR_ERR_NOMEM;
  v->ctx = ctx;
  v->onx = v->nx = nx; v->ony = v->ny = ny; v->onz = v->nz = nz;
  const size_t bytes = v->count() * sizeof(int16_t);
  cudaError_t e = pool_alloc(ctx, &v->original, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(v->original, voxels, bytes, cudaMemcpyHostToDevice, ctx->stream);
  v->zlo = zlo; v->zhi = zhi;
  int s = VR_OK;
  if (e != cudaSuccess) { vr_set_error("vr_volume_upload: %s", cudaGetErrorString(e)); s = VR_ERR_CUDA; }
  if (s == VR_OK) s = vrk_fetch_stats(ctx, v->original, nx, ny, nz, v->stats, zlo, zhi);  // reference_volume.cpp:22-41
  if (s != VR_OK) {
    cudaStreamSynchronize(ctx->stream);  // the (possibly staged) copy must not outlive the caller's buffer
    pool_free(ctx, v->original);
    delete v;
    return s;
  }
  *out = v;
  return VR_OK;
}

extern "C" int vr_volume_upload(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz, vr_volume** out) {
  return volume_upload_impl(ctx, voxels, nx, ny, nz, 0, nz, out);
}

// a volume that already lives on this device (decoded or generated there): device-to-device copy, then fetch_stats
extern "C" int vr_volume_upload_device(vr_ctx* ctx, const int16_t* device_voxels, int nx, int ny, int nz, vr_volume** out) {
  VR_REQUIRE(ctx && device_voxels && out, "vr_volume_upload_device: null argument");
  VR_REQUIRE(nx > 0 && ny > 0 && nz > 0, "vr_volume_upload_device: dimensions must be positive");
  VR_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 32) - 1, "vr_volume_upload_device: more than 2^32-2 voxels");
  VR_CUDA(cudaSetDevice(ctx->device));
  cudaPointerAttributes at{};
  VR_REQUIRE(cudaPointerGetAttributes(&at, device_voxels) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device == ctx->device,
             "vr_volume_upload_device: not a pointer to memory of the context's device");
  vr_volume* v = new (std::nothrow) vr_volume();
  if (!v) return VR_ERR_NOMEM;
  v->ctx = ctx;
  v->onx = v->nx = nx; v->ony = v->ny = ny; v->onz = v->nz = nz;
  v->zlo = 0; v->zhi = nz;
  const size_t bytes = v->count() * sizeof(int16_t);
  cudaError_t e = pool_alloc(ctx, &v->original, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(v->original, device_voxels, bytes, cudaMemcpyDeviceToDevice, ctx->stream);
  int s = VR_OK;
  if (e != cudaSuccess) { vr_set_error("vr_volume_upload_device: %s", cudaGetErrorString(e)); s = VR_ERR_CUDA; }
  if (s == VR_OK) s = vrk_fetch_stats(ctx, v->original, nx, ny, nz, v->stats, 0, nz);
  if (s != VR_OK) { cudaStreamSynchronize(ctx->stream); pool_free(ctx, v->original); delete v; return s; }
  *out = v;
  return VR_OK;
}

// Asynchronous ingest (no reference counterpart: clw_image pushes are blocking, clw_image.hpp:206).  Returns at once; the copy
// from (preferably pinned) host memory and fetch_stats run on the context's copy stream beside whatever the compute stream
// is doing, e.g. the previous job's SDF build and frames.  `voxels` must stay valid and unchanged until vr_volume_wait (or
// any other call that uses the volume, which waits implicitly) returns.
extern "C" int vr_volume_upload_async(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz, vr_volume** out) {
  VR_REQUIRE(ctx && voxels && out, "vr_volume_upload_async: null argument");
  VR_REQUIRE(nx > 0 && ny > 0 && nz > 0, "vr_volume_upload_async: dimensions must be positive");
  VR_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 32) - 1, "vr_volume_upload_async: more than 2^32-2 voxels");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_volume* v = new (std::nothrow) vr_volume();
  if (!v) return VR_ERR_NOMEM;
  v->ctx = ctx;
  v->onx = v->nx = nx; v->ony = v->ny = ny; v->onz = v->nz = nz;
  v->zlo = 0; v->zhi = nz;
  const size_t bytes = v->count() * sizeof(int16_t);
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&v->original), bytes, ctx->copy_stream);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&v->stats_dev), 4 * sizeof(int32_t), ctx->copy_stream);
  if (e == cudaSuccess) e = pinned_acquire(ctx, reinterpret_cast<void**>(&v->stats_pin), 64);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&v->ready, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMemcpyAsync(v->original, voxels, bytes, cudaMemcpyHostToDevice, ctx->copy_stream);
  int s = VR_OK;
  if (e != cudaSuccess) { vr_set_error("vr_volume_upload_async: %s", cudaGetErrorString(e)); s = VR_ERR_CUDA; }
  if (s == VR_OK) s = vrk_fetch_stats_enqueue(ctx, ctx->copy_stream, v->original, nx, ny, nz, 0, nz, v->stats_dev, v->stats_pin);
  if (s == VR_OK && (e = cudaEventRecord(v->ready, ctx->copy_stream)) != cudaSuccess) {
    vr_set_error("vr_volume_upload_async: %s", cudaGetErrorString(e));
    s = VR_ERR_CUDA;
  }
  if (s != VR_OK) {  // undo whatever was set up; the copy stream may still hold work that touches the buffers
    cudaStreamSynchronize(ctx->copy_stream);
    if (v->original) cudaFreeAsync(v->original, ctx->copy_stream);
    if (v->stats_dev) cudaFreeAsync(v->stats_dev, ctx->copy_stream);
    if (v->stats_pin) pinned_release(ctx, v->stats_pin);
    if (v->ready) cudaEventDestroy(v->ready);
    delete v;
    return s;
  }
  v->pending = true;
  *out = v;
  return VR_OK;
}

// VR_SAMPLING_HW_LINEAR for the volume kernels: the box-averaged copy of the current volume (vr_volume_ops.cu k_boxavg)
static void volume_release_textures(vr_volume* v) {
  if (!v->box) return;
  pool_free(v->ctx, v->box);
  v->box = nullptr; v->box_px = 0;
}
static int volume_build_textures(vr_volume* v) {
  volume_release_textures(v);
  v->box_px = (v->nx + 1 + 7) / 8 * 8;
  const size_t n = (size_t)v->box_px * (v->ny + 1) * (v->nz + 1);
  VR_REQUIRE(n < ((size_t)1 << 32) - 1, "hw-linear sampling: the box-averaged volume would exceed 2^32-2 voxels");
  VR_CUDA(pool_alloc(v->ctx, &v->box, n * sizeof(int16_t)));
  int st = vrk_boxavg(v->ctx, v->current(), v->nx, v->ny, v->nz, v->box, v->box_px);
  if (st != VR_OK) volume_release_textures(v);
  return st;
}

extern "C" int vr_volume_set_sampling(vr_volume* v, int mode) {
  VR_REQUIRE(v && (mode == VR_SAMPLING_NEAREST || mode == VR_SAMPLING_HW_LINEAR), "vr_volume_set_sampling: unknown mode");
  VR_TRY(volume_finish(v));
  if (mode == v->sampling) return VR_OK;
  VR_CUDA(cudaSetDevice(v->ctx->device));
  if (mode == VR_SAMPLING_HW_LINEAR) {
    v->raw_range[0] = v->stats[0]; v->raw_range[1] = v->stats[1];  // the NEAREST extremes: clip and filter can only narrow them
    VR_TRY(volume_build_textures(v));
    int st = vrk_fetch_stats_linear(v->ctx, v->box, v->box_px, v->current(), v->nx, v->ny, v->nz, v->stats, v->zlo, v->zhi);
    if (st != VR_OK) { volume_release_textures(v); return st; }
  } else {
    volume_release_textures(v);
    VR_TRY(vrk_fetch_stats(v->ctx, v->current(), v->nx, v->ny, v->nz, v->stats, v->zlo, v->zhi));
  }
  v->sampling = mode;
  return VR_OK;
}

extern "C" int vr_volume_wait(vr_volume* v) {
  VR_REQUIRE(v, "vr_volume_wait: null argument");
  return volume_finish(v);
}

// z-slab of a larger volume (multi-GPU sharding, no reference counterpart): planes [z_lo, z_hi) are this rank's, the planes
// around them are halo — read by gradient taps, filter taps and the SDF wave, not counted by stats / histogram
extern "C" int vr_volume_upload_slab(vr_ctx* ctx, const int16_t* voxels, int nx, int ny, int nz_ext, int z_lo, int z_hi,
                                     vr_volume** out) {
  return volume_upload_impl(ctx, voxels, nx, ny, nz_ext, z_lo, z_hi, out);
}

extern "C" int vr_volume_download_planes(const vr_volume* v, int z0, int nplanes, int16_t* out) {
  VR_REQUIRE(v && out && z0 >= 0 && nplanes > 0 && z0 + nplanes <= v->nz, "vr_volume_download_planes: bad argument");
  VR_TRY(volume_finish(v));
  VR_CUDA(cudaSetDevice(v->ctx->device));
  const size_t plane = (size_t)v->nx * v->ny;
  VR_CUDA(cudaMemcpyAsync(out, v->current() + plane * z0, plane * nplanes * sizeof(int16_t), cudaMemcpyDeviceToHost,
                          v->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return VR_OK;
}

extern "C" void vr_volume_destroy(vr_volume* v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  volume_finish(v);
  cudaStreamSynchronize(v->ctx->stream);
  volume_release_textures(v);
  pool_free(v->ctx, v->original);
  pool_free(v->ctx, v->cropped);
  cudaStreamSynchronize(v->ctx->stream);
  delete v;
}

extern "C" int vr_volume_stats(const vr_volume* v, int32_t out[4]) {
  VR_REQUIRE(v && out, "vr_volume_stats: null argument");
  VR_TRY(volume_finish(v));
  memcpy(out, v->stats, sizeof(v->stats));
  return VR_OK;
}

extern "C" int vr_volume_dims(const vr_volume* v, int out[3]) {
  VR_REQUIRE(v && out, "vr_volume_dims: null argument");
  out[0] = v->nx; out[1] = v->ny; out[2] = v->nz;
  return VR_OK;
}

extern "C" int vr_volume_clip(vr_volume* v, const uint32_t mn[3], const uint32_t mx[3]) {
  VR_REQUIRE(v && mn && mx, "vr_volume_clip: null argument");
  VR_TRY(volume_finish(v));
  // reference_volume.cpp:57-59 asserts min < max; reads beyond the original are border reads (0)
  VR_REQUIRE(mn[0] < mx[0] && mn[1] < mx[1] && mn[2] < mx[2], "vr_volume_clip: min must be < max on every axis");
  VR_CUDA(cudaSetDevice(v->ctx->device));
  const int nx = (int)(mx[0] - mn[0]), ny = (int)(mx[1] - mn[1]), nz = (int)(mx[2] - mn[2]);
  int16_t* dst = nullptr;
  VR_CUDA(pool_alloc(v->ctx, &dst, (size_t)nx * ny * nz * sizeof(int16_t)));
  int s = vrk_clip(v->ctx, v->original, v->onx, v->ony, v->onz, mn, dst, nx, ny, nz);
  if (s != VR_OK) { pool_free(v->ctx, dst); return s; }
  VR_CUDA(cudaStreamSynchronize(v->ctx->stream));
  pool_free(v->ctx, v->cropped);
  v->cropped = dst;
  v->nx = nx; v->ny = ny; v->nz = nz;
  v->zlo = 0; v->zhi = nz;
  v->generation++;
  if (v->sampling == VR_SAMPLING_HW_LINEAR) VR_TRY(volume_build_textures(v));  // the textures follow the current volume
  return VR_OK;
}

extern "C" int vr_volume_filter(vr_volume* v) {
  VR_REQUIRE(v, "vr_volume_filter: null argument");
  VR_TRY(volume_finish(v));
  VR_CUDA(cudaSetDevice(v->ctx->device));
  int16_t* dst = nullptr;
  VR_CUDA(pool_alloc(v->ctx, &dst, v->count() * sizeof(int16_t)));
  int s = v->sampling == VR_SAMPLING_HW_LINEAR ? vrk_bilateral_linear(v->ctx, v->box, v->box_px, v->nx, v->ny, v->nz, dst)
                                               : vrk_bilateral(v->ctx, v->current(), dst, v->nx, v->ny, v->nz);
  if (s != VR_OK) { pool_free(v->ctx, dst); return s; }
  VR_CUDA(cudaStreamSynchronize(v->ctx->stream));
  // `ref = std::move(buffer)` (reference_volume.cpp:77): the filtered data replaces the current volume
  if (v->cropped) { pool_free(v->ctx, v->cropped); v->cropped = dst; }
  else { pool_free(v->ctx, v->original); v->original = dst; }
  v->generation++;
  if (v->sampling == VR_SAMPLING_HW_LINEAR) VR_TRY(volume_build_textures(v));
  return VR_OK;
}

extern "C" int vr_volume_download(const vr_volume* v, int16_t* out) {
  VR_REQUIRE(v && out, "vr_volume_download: null argument");
  VR_TRY(volume_finish(v));
  VR_CUDA(cudaSetDevice(v->ctx->device));
  VR_CUDA(cudaMemcpyAsync(out, v->current(), v->count() * sizeof(int16_t), cudaMemcpyDeviceToHost, v->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(v->ctx->stream));
  return VR_OK;
}

extern "C" int vr_histogram(const vr_volume* v, int width, int height, const float range[4], uint32_t* bins_out) {
  VR_REQUIRE(v && range && bins_out, "vr_histogram: null argument");
  VR_TRY(volume_finish(v));
  VR_REQUIRE(width > 0 && height > 0 && (size_t)width * height < ((size_t)1 << 31), "vr_histogram: bad bin grid");
  VR_CUDA(cudaSetDevice(v->ctx->device));
  uint32_t* bins = nullptr;
  const size_t bytes = sizeof(uint32_t) * (size_t)width * height;
  VR_CUDA(pool_alloc(v->ctx, &bins, bytes));
  int s = v->sampling == VR_SAMPLING_HW_LINEAR
              ? vrk_histogram_linear(v->ctx, v->box, v->box_px, v->current(), v->nx, v->ny, v->nz, width, height, range, bins, v->zlo, v->zhi,
                                     v->stats[0], v->raw_range[0], v->raw_range[1])
              : vrk_histogram(v->ctx, v->current(), v->nx, v->ny, v->nz, width, height, range, bins, v->zlo, v->zhi, v->stats[0], v->stats[1]);
  if (s == VR_OK) {
    cudaError_t e = cudaMemcpyAsync(bins_out, bins, bytes, cudaMemcpyDeviceToHost, v->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(v->ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_histogram: %s", cudaGetErrorString(e)); s = VR_ERR_CUDA; }
  }
  pool_free(v->ctx, bins);
  return s;
}

// ---- environment map -------------------------------------------------------------------------------------------
extern "C" int vr_envmap_bind(vr_ctx* ctx, const uint8_t* rgba8, int w, int h, vr_envmap** out) {
  VR_REQUIRE(ctx && rgba8 && out, "vr_envmap_bind: null argument");
  VR_REQUIRE(w > 0 && h > 0, "vr_envmap_bind: dimensions must be positive");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_envmap* e = new (std::nothrow) vr_envmap();
  if (!e) return VR_ERR_NOMEM;
  e->ctx = ctx; e->w = w; e->h = h;
  cudaError_t ce = pool_alloc(ctx, &e->texels, (size_t)w * h * 4);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->texels, rgba8, (size_t)w * h * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
  if (ce != cudaSuccess) {
    vr_set_error("vr_envmap_bind: %s", cudaGetErrorString(ce));
    pool_free(ctx, e->texels);
    delete e;
    return VR_ERR_CUDA;
  }
  *out = e;
  return VR_OK;
}

extern "C" void vr_envmap_destroy(vr_envmap* e) {
  if (!e) return;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(e->ctx->stream);
  pool_free(e->ctx, e->texels);
  cudaStreamSynchronize(e->ctx->stream);
  delete e;
}

// ---- SDF -----------------------------------------------------------------------------------------------------
int sdf_build_impl(vr_ctx* ctx, const vr_volume* vol, const TfTable& tf, vr_sdf** out, int sharded) {
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_sdf* s = new (std::nothrow) vr_sdf();
  if (!s) return VR_ERR_NOMEM;
  s->ctx = ctx; s->nx = vol->nx; s->ny = vol->ny; s->nz = vol->nz;
  cudaError_t e = pool_alloc(ctx, &s->field, vrk_sdf_field_bytes(vol->nx, vol->ny, vol->nz));
  if (e != cudaSuccess) { delete s; vr_set_error("vr_sdf_build: %s", cudaGetErrorString(e)); return VR_ERR_CUDA; }
  // the same values in a 3-D array behind a surface object: what k_trace_pt gathers from
  if (array3d_acquire(ctx, vol->nx, vol->ny, vol->nz, 8, &s->arr, &s->surf) != VR_OK) {
    pool_free(ctx, s->field);
    delete s;
    return VR_ERR_CUDA;
  }
  int st = sharded ? vrk_sdf_build_sharded(ctx, vol->current(), vol->nx, vol->ny, vol->nz, tf, s->field, &s->levels, &s->max_it, s->surf, sharded == 1)
                   : vrk_sdf_build(ctx, vol->current(), vol->nx, vol->ny, vol->nz, tf, s->field, &s->levels, &s->max_it, s->surf);
  if (st != VR_OK) { vr_sdf_destroy(s); return st; }
  *out = s;
  return VR_OK;
}

// tmp_color of a clause is (int)(c*255) with c in [0,1] (tf_part.cpp:60-77): the packed accumulators of k_trace_pt rely on it
static int tf_check_colours(const vr_tf_rect* rects, int n, const char* who) {
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 4; ++k)
      if (rects[i].rgba[k] < 0 || rects[i].rgba[k] > 255) {
        vr_set_error("%s: clause %d: colour component %d = %d is outside [0,255]", who, i, k, rects[i].rgba[k]);
        return VR_ERR_INVALID;
      }
  return VR_OK;
}

extern "C" int vr_sdf_build(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out) {
  VR_REQUIRE(ctx && vol && out && (rects || n_rects == 0), "vr_sdf_build: null argument");
  VR_TRY(volume_finish(vol));
  VR_REQUIRE(n_rects >= 0 && n_rects <= VR_TF_MAX_RECTS, "vr_sdf_build: too many TF clauses");
  VR_TRY(tf_check_colours(rects, n_rects, "vr_sdf_build"));
  return sdf_build_impl(ctx, vol, vr_make_tf_table(rects, n_rects), out, 0);
}

extern "C" void vr_sdf_destroy(vr_sdf* s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  pool_free(s->ctx, s->field);
  cudaStreamSynchronize(s->ctx->stream);
  array3d_release(s->ctx, s->arr);  // the array goes back to the context's cache
  delete s;
}

extern "C" int vr_sdf_download(const vr_sdf* s, int8_t* out) {
  VR_REQUIRE(s && out, "vr_sdf_download: null argument");
  VR_CUDA(cudaSetDevice(s->ctx->device));
  const size_t n = (size_t)s->nx * s->ny * s->nz;
  int8_t* linear = nullptr;
  VR_CUDA(pool_alloc(s->ctx, &linear, n));
  int st = vrk_sdf_unbrick(s->ctx, s->field, s->nx, s->ny, s->nz, linear);
  if (st == VR_OK) {
    cudaError_t e = cudaMemcpyAsync(out, linear, n, cudaMemcpyDeviceToHost, s->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_sdf_download: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(s->ctx, linear);
  return st;
}

extern "C" int vr_sdf_levels(const vr_sdf* s) { return s ? s->levels : 0; }

extern "C" int vr_sdf_checksum(const vr_sdf* s, uint64_t* out) {
  VR_REQUIRE(s && out, "vr_sdf_checksum: null argument");
  VR_CUDA(cudaSetDevice(s->ctx->device));
  return vrk_checksum(s->ctx, s->field, vrk_sdf_field_bytes(s->nx, s->ny, s->nz), out);
}
extern "C" const int16_t* vr_volume_device_ptr(const vr_volume* v) {
  if (!v || volume_finish(v) != VR_OK) return nullptr;
  return v->current();
}
extern "C" int vr_volume_checksum(const vr_volume* v, uint64_t* out) {
  VR_REQUIRE(v && out, "vr_volume_checksum: null argument");
  VR_TRY(volume_finish(v));
  VR_CUDA(cudaSetDevice(v->ctx->device));
  return vrk_checksum(v->ctx, v->current(), v->count() * sizeof(int16_t), out);
}

// ---- z-slab SDF build (multi-GPU sharding; driver: cl_volume_renderer_b200/parallel.py) -------------------------------------
extern "C" int vr_sdf_slab_create(vr_ctx* ctx, const vr_volume* ext_slab, const vr_tf_rect* rects, int n_rects, int max_it_global,
                                  vr_sdf_slab** out) {
  VR_REQUIRE(ctx && ext_slab && out && (rects || n_rects == 0), "vr_sdf_slab_create: null argument");
  VR_TRY(volume_finish(ext_slab));
  VR_REQUIRE(n_rects >= 0 && n_rects <= VR_TF_MAX_RECTS, "vr_sdf_slab_create: too many TF clauses");
  VR_REQUIRE(max_it_global >= 1 && max_it_global <= 127, "vr_sdf_slab_create: max_it must be in [1,127]");
  VR_CUDA(cudaSetDevice(ctx->device));
  return vrk_sdf_slab_create(ctx, ext_slab->current(), ext_slab->nx, ext_slab->ny, ext_slab->nz, vr_make_tf_table(rects, n_rects),
                             max_it_global, out);
}
extern "C" int vr_sdf_slab_advance(vr_sdf_slab* s, int nlevels, int* levels_done) {
  VR_REQUIRE(s && nlevels >= 0, "vr_sdf_slab_advance: bad argument");
  return vrk_sdf_slab_advance(s, nlevels, levels_done);
}
extern "C" void* vr_sdf_slab_bits(vr_sdf_slab* s) { return s ? (void*)vrk_sdf_slab_bits(s) : nullptr; }
extern "C" size_t vr_sdf_slab_plane_words(const vr_sdf_slab* s) { return s ? vrk_sdf_slab_plane_words(s) : 0; }
extern "C" int vr_sdf_slab_mark_imported(vr_sdf_slab* s) {
  VR_REQUIRE(s, "vr_sdf_slab_mark_imported: null argument");
  vrk_sdf_slab_mark_imported(s);
  return VR_OK;
}
extern "C" int vr_sdf_slab_finished(const vr_sdf_slab* s) { return s ? (vrk_sdf_slab_finished(s) ? 1 : 0) : 1; }
extern "C" int vr_sdf_slab_download(vr_sdf_slab* s, vr_ctx* ctx, int nx, int ny, int nz_ext, int z0, int nplanes, int8_t* out) {
  VR_REQUIRE(s && ctx && out && z0 >= 0 && nplanes > 0 && z0 + nplanes <= nz_ext, "vr_sdf_slab_download: bad argument");
  VR_CUDA(cudaSetDevice(ctx->device));
  int8_t *field = nullptr, *linear = nullptr;
  const size_t n = (size_t)nx * ny * nz_ext;
  VR_CUDA(pool_alloc(ctx, &field, vrk_sdf_field_bytes(nx, ny, nz_ext)));
  VR_CUDA(pool_alloc(ctx, &linear, n));
  int st = vrk_sdf_slab_assemble(s, field);
  if (st == VR_OK) st = vrk_sdf_unbrick(ctx, field, nx, ny, nz_ext, linear);
  if (st == VR_OK) {
    const size_t plane = (size_t)nx * ny;
    cudaError_t e = cudaMemcpyAsync(out, linear + plane * z0, plane * nplanes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_sdf_slab_download: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, linear);
  pool_free(ctx, field);
  return st;
}
extern "C" void vr_sdf_slab_destroy(vr_sdf_slab* s) { vrk_sdf_slab_destroy(s); }

// ---- renderer --------------------------------------------------------------------------------------------------
extern "C" int vr_renderer_create(vr_ctx* ctx, int width, int height, vr_renderer** out) {
  VR_REQUIRE(ctx && out, "vr_renderer_create: null argument");
  VR_REQUIRE(width > 0 && height > 0, "vr_renderer_create: frame size must be positive");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_renderer* r = new (std::nothrow) vr_renderer();
  if (!r) return VR_ERR_NOMEM;
  r->ctx = ctx; r->W = width; r->H = height; r->row0 = 0; r->row1 = height;
  const size_t px = (size_t)width * height;
  cudaError_t e = pool_alloc(ctx, &r->frame, px * 4);
  if (e == cudaSuccess) e = pool_alloc(ctx, &r->hit, px * 4);
  if (e == cudaSuccess) e = pool_alloc(ctx, &r->counters, 8 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = pool_alloc(ctx, &r->bbox_dev, 4 * sizeof(int));
  if (e == cudaSuccess) e = pinned_acquire(ctx, reinterpret_cast<void**>(&r->bbox_pin), 64);
  if (e == cudaSuccess) e = pinned_acquire(ctx, reinterpret_cast<void**>(&r->frame_host), px * 4);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->frame, 0, px * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->hit, 0xFF, px * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(r->counters, 0, 8 * sizeof(unsigned long long), ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    vr_set_error("vr_renderer_create: %s", cudaGetErrorString(e));
    vr_renderer_destroy(r);  // releases whatever was acquired
    return VR_ERR_CUDA;
  }
  *out = r;
  return VR_OK;
}

// VR_SAMPLING_HW_LINEAR: the current volume and the environment map behind texture objects (CUDA arrays owned by the renderer)
static void release_textures(vr_renderer* r) {
  if (!r->vol_tex && !r->env_tex && !r->vol_arr && !r->env_arr && !r->lin_arr) return;
  cudaStreamSynchronize(r->ctx->stream);
  array3d_release(r->ctx, r->lin_arr);
  r->lin_arr = nullptr; r->lin_surf = 0;
  if (r->vol_tex) cudaDestroyTextureObject(r->vol_tex);
  if (r->env_tex) cudaDestroyTextureObject(r->env_tex);
  array3d_release(r->ctx, r->vol_arr);  // back to the context's cache: a renderer created for the next job of the same size reuses it
  if (r->env_arr) cudaFreeArray(r->env_arr);
  r->vol_tex = r->env_tex = 0;
  r->vol_arr = r->env_arr = nullptr;
  r->tex_dims[0] = 0;
}

// The arrays are kept while the sizes stay the same (a flush for a new volume of the same size, a TF edit): cudaMalloc3DArray /
// cudaFreeArray synchronise the device and cost milliseconds; only the contents are copied again.
static int build_textures(vr_renderer* r) {
  vr_ctx* ctx = r->ctx;
  const vr_volume* v = r->vol;
  const vr_envmap* env = r->env;
  const bool same = r->vol_tex && r->env_tex && r->lin_arr && r->tex_dims[0] == v->nx && r->tex_dims[1] == v->ny && r->tex_dims[2] == v->nz &&
                    r->tex_dims[3] == env->w && r->tex_dims[4] == env->h;
  if (!same) release_textures(r);
  int st = VR_OK;
  cudaError_t e = cudaSuccess;
  if (!same) {
    cudaSurfaceObject_t none = 0;
    if (array3d_acquire(ctx, v->nx, v->ny, v->nz, 17, &r->vol_arr, &none) != VR_OK) { release_textures(r); return VR_ERR_CUDA; }
  }
  if (e == cudaSuccess) {
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(const_cast<int16_t*>(v->current()), (size_t)v->nx * sizeof(int16_t), v->nx, v->ny);
    p.dstArray = r->vol_arr;
    p.extent = make_cudaExtent(v->nx, v->ny, v->nz);
    p.kind = cudaMemcpyDeviceToDevice;
    e = cudaMemcpy3DAsync(&p, ctx->stream);
  }
  if (e == cudaSuccess && !same) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = r->vol_arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;  // CLK_ADDRESS_CLAMP: border colour 0
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    e = cudaCreateTextureObject(&r->vol_tex, &rd, &td, nullptr);
  }
  if (e == cudaSuccess && !same) {
    cudaChannelFormatDesc d8 = cudaCreateChannelDesc<uchar4>();
    e = cudaMallocArray(&r->env_arr, &d8, env->w, env->h);
  }
  if (e == cudaSuccess)
    e = cudaMemcpy2DToArrayAsync(r->env_arr, 0, 0, env->texels, (size_t)env->w * 4, (size_t)env->w * 4, env->h, cudaMemcpyDeviceToDevice,
                                 ctx->stream);
  if (e == cudaSuccess && !same) {
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = r->env_arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;  // CLK_ADDRESS_CLAMP_TO_EDGE
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;                                       // CLK_NORMALIZED_COORDS_TRUE
    e = cudaCreateTextureObject(&r->env_tex, &rd, &td, nullptr);
  }
  if (e != cudaSuccess) {
    vr_set_error("vr_renderer_flush: texture setup for hw-linear sampling: %s", cudaGetErrorString(e));
    release_textures(r);
    return VR_ERR_CUDA;
  }
  // the step field: SDF byte + quiet-octant bits per voxel cell (vr_quiet.cu), from the SDF the flush has just built
  if (!same) st = array3d_acquire(ctx, v->nx, v->ny, v->nz, 16, &r->lin_arr, &r->lin_surf);
  if (st == VR_OK) st = vrk_lin_field_build(ctx, v->current(), v->nx, v->ny, v->nz, r->sdf->field, r->tf_active, r->lin_surf);
  if (st != VR_OK) { release_textures(r); return st; }
  r->tex_dims[0] = v->nx; r->tex_dims[1] = v->ny; r->tex_dims[2] = v->nz; r->tex_dims[3] = env->w; r->tex_dims[4] = env->h;
  return VR_OK;
}

extern "C" int vr_renderer_set_sampling(vr_renderer* r, int mode) {
  VR_REQUIRE(r && (mode == VR_SAMPLING_NEAREST || mode == VR_SAMPLING_HW_LINEAR), "vr_renderer_set_sampling: unknown mode");
  if (mode != r->sampling) {
    VR_CUDA(cudaSetDevice(r->ctx->device));
    release_textures(r);  // rebuilt by the next flush when needed
    r->sampling = mode;
    r->primary_valid = false;
  }
  return VR_OK;
}

extern "C" void vr_renderer_destroy(vr_renderer* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  release_textures(r);
  vr_sdf_destroy(r->sdf);
  pool_free(r->ctx, r->cache);
  pool_free(r->ctx, r->hit);
  pool_free(r->ctx, r->frame);
  pool_free(r->ctx, r->counters);
  pool_free(r->ctx, r->xchg);
  pool_free(r->ctx, r->queue);
  pool_free(r->ctx, r->filtered);
  pool_free(r->ctx, r->xc_idx);
  pool_free(r->ctx, r->xc_counts);
  pool_free(r->ctx, r->gather_buf);
  cudaStreamSynchronize(r->ctx->stream);
  for (cudaEvent_t e : r->ev) cudaEventDestroy(e);
  pool_free(r->ctx, r->bbox_dev);
  cudaStreamSynchronize(r->ctx->stream);
  pinned_release(r->ctx, r->frame_host);
  pinned_release(r->ctx, r->bbox_pin);
  delete r;
}

extern "C" int vr_renderer_set_scene(vr_renderer* r, const vr_volume* vol, const vr_envmap* env) {
  VR_REQUIRE(r && vol && env, "vr_renderer_set_scene: null argument");
  VR_REQUIRE(vol->ctx == r->ctx && env->ctx == r->ctx, "vr_renderer_set_scene: objects belong to another context");
  r->vol = vol;
  r->env = env;
  r->primary_valid = false;
  return VR_OK;
}

extern "C" int vr_renderer_set_tf(vr_renderer* r, const vr_tf_rect* rects, int n_rects) {
  VR_REQUIRE(r && (rects || n_rects == 0), "vr_renderer_set_tf: null argument");
  VR_REQUIRE(n_rects >= 0 && n_rects <= VR_TF_MAX_RECTS, "vr_renderer_set_tf: too many TF clauses");
  VR_TRY(tf_check_colours(rects, n_rects, "vr_renderer_set_tf"));
  r->tf_pending = vr_make_tf_table(rects, n_rects);
  r->have_tf = true;
  return VR_OK;
}

extern "C" int vr_renderer_set_tf_code(vr_renderer* r, const char* src) {
  VR_REQUIRE(r && src, "vr_renderer_set_tf_code: null argument");
  vr_tf_rect rects[VR_TF_MAX_RECTS];
  int n = 0;
  VR_TRY(vr_tf_parse(src, rects, VR_TF_MAX_RECTS, &n));
  return vr_renderer_set_tf(r, rects, n);
}

extern "C" int vr_renderer_reset_cache(vr_renderer* r) {
  VR_REQUIRE(r && r->cache, "vr_renderer_reset_cache: no cache (call vr_renderer_flush first)");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  if (r->cache_exposed) r->cache_dirty = 2;
  int st = VR_OK;
  if (r->cache_dirty == 1) st = vrk_cache_reset_hits(r->ctx, r->cache, r->hit, (size_t)r->W * r->H);
  else if (r->cache_dirty == 2) st = vrk_cache_reset(r->ctx, r->cache, r->cache_voxels);
  if (st == VR_OK) r->cache_dirty = 0;
  return st;
}

// same event predicate: clause count, kinds, value and gradient ranges in order — everything of is_event_gen except the colours
static bool tf_same_predicate(const TfTable& a, const TfTable& b) {
  if (a.n != b.n) return false;
  for (int i = 0; i < a.n; ++i) {
    const vr_tf_rect &x = a.r[i], &y = b.r[i];
    if (x.flags != y.flags || memcmp(&x.min_v, &y.min_v, sizeof(float)) || memcmp(&x.max_v, &y.max_v, sizeof(float))) return false;
    if ((x.flags & VR_TF_USE_GRADIENT) && (memcmp(&x.min_g, &y.min_g, sizeof(float)) || memcmp(&x.max_g, &y.max_g, sizeof(float)))) return false;
  }
  return true;
}

extern "C" int vr_renderer_flush(vr_renderer* r) {
  VR_REQUIRE(r && r->vol && r->env, "vr_renderer_flush: no scene bound (vr_renderer_set_scene)");
  VR_TRY(volume_finish(r->vol));
  VR_REQUIRE(r->have_tf, "vr_renderer_flush: no transfer function set");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  // renderer.cpp:29-30 — reallocate the cache only when the volume size changed
  const size_t voxels = r->vol->count();
  bool fresh_cache = false, forked = false;
  if (voxels != r->cache_voxels) {
    fresh_cache = true;
    VR_CUDA(cudaStreamSynchronize(r->ctx->stream));
    pool_free(r->ctx, r->cache);
    r->cache = nullptr;
    r->cache_voxels = 0;
    VR_CUDA(pool_alloc(r->ctx, &r->cache, voxels * 8));
    r->cache_voxels = voxels;
  }
  r->primary_valid = false;
  // renderer.cpp:32-35 (buffer_reset) and :42 (new signed_distance_field) do not depend on each other: the reset is pure store
  // bandwidth, the SDF build a chain of 125 latency-bound levels, so the reset runs on a second stream beside the build.
  // A cache that was not reallocated and has seen one camera only is cleared through the hit buffer (vr_renderer_reset_cache).
  vr_ctx* ctx = r->ctx;
  if (!fresh_cache && !r->cache_exposed && r->cache_dirty <= 1) {
    if (r->cache_dirty == 1) VR_TRY(vrk_cache_reset_hits(ctx, r->cache, r->hit, (size_t)r->W * r->H));
  } else {
    VR_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
    VR_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    VR_TRY(vrk_cache_reset(ctx, r->cache, r->cache_voxels, ctx->aux_stream));
    VR_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    forked = true;
  }
  r->cache_dirty = 0;
  // Incremental rebuild: the SDF and the step field depend on the transfer function only through its event
  // PREDICATE (value / gradient ranges, clause order and kind), not through the colours.  A flush after a colour edit — the
  // common edit in the reference's UI, which re-JITs three kernels and rebuilds the field for it (renderer.cpp:39-42) — keeps
  // both when volume, environment map and sampling are those of the last flush.
  const bool keep_field = r->sdf && r->flushed_vol == r->vol && r->flushed_env == r->env && r->flushed_generation == r->vol->generation &&
                          r->flushed_sampling == r->sampling && r->flushed_sharded == r->sharded_build &&
                          tf_same_predicate(r->tf_active, r->tf_pending) &&
                          (r->sampling != VR_SAMPLING_HW_LINEAR || (r->vol_tex && r->env_tex && r->lin_surf));
  r->tf_active = r->tf_pending;                                  // renderer.cpp:39
  if (!keep_field) {
    vr_sdf* fresh = nullptr;                                     // renderer.cpp:42
    // z-slab sharded only where slabs pay: a volume whose levels are a single wave of CTAs is built on every rank (no collectives)
    const bool slabs = r->sharded_build && !vrk_sdf_single_wave(ctx, r->vol->nx, r->vol->ny, r->vol->nz);
    int st = sdf_build_impl(ctx, r->vol, r->tf_active, &fresh, slabs ? 1 : 0);
    if (forked) VR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // later work on the compute stream sees the reset cache
    forked = false;
    VR_TRY(st);
    vr_sdf_destroy(r->sdf);
    r->sdf = fresh;
  }
  if (forked) VR_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  r->fields_kept = keep_field;
  VR_CUDA(cudaMemsetAsync(r->hit, 0xFF, (size_t)r->W * r->H * 4, r->ctx->stream));
  r->flushed_vol = nullptr;
  if (r->sampling == VR_SAMPLING_HW_LINEAR && !keep_field) VR_TRY(build_textures(r));
  r->flushed_sampling = r->sampling; r->flushed_sharded = r->sharded_build;
  r->flushed_vol = r->vol; r->flushed_env = r->env; r->flushed_generation = r->vol->generation;
  r->flush_count++;
  return VR_OK;
}

// SDF, cache, hit buffer and textures belong to the scene of the last flush: tracing or resolving with another volume bound, or
// with the bound volume clipped / filtered since, would index them with the wrong dimensions
static int check_flushed(const vr_renderer* r, const char* who) {
  if (!r->sdf || !r->cache) { vr_set_error("%s: call vr_renderer_flush first", who); return VR_ERR_INVALID; }
  if (r->vol != r->flushed_vol || r->env != r->flushed_env || r->vol->generation != r->flushed_generation) {
    vr_set_error("%s: the scene changed since the last vr_renderer_flush (set_scene, vr_volume_clip or vr_volume_filter): flush required", who);
    return VR_ERR_INVALID;
  }
  return VR_OK;
}

// The blocking pull of renderer.cpp:150.  Into the renderer-owned host frame the pull is incremental: while the primary records
// of the last complete pull are still the current ones (same camera, rows and scene; vr_renderer_set_primary_reuse(r, 2)), only
// shaded pixels can have changed, and they lie inside the bounding box k_primary reduced.
static int read_frame(vr_renderer* r, uint8_t* host_rgba) {
  if (!host_rgba) return VR_OK;
  vr_ctx* ctx = r->ctx;
  const bool own = host_rgba == r->frame_host;
  if (own && r->host_epoch != 0 && r->host_epoch == r->primary_epoch && r->primary_valid) {
    const int x0 = r->bbox_pin[0], y0 = r->bbox_pin[1], x1 = r->bbox_pin[2], y1 = r->bbox_pin[3];
    if (x1 >= x0 && y1 >= y0) {
      const size_t pitch = (size_t)r->W * 4, off = ((size_t)y0 * r->W + x0) * 4;
      VR_CUDA(cudaMemcpy2DAsync(host_rgba + off, pitch, reinterpret_cast<const uint8_t*>(r->frame) + off, pitch, (size_t)(x1 - x0 + 1) * 4,
                                (size_t)(y1 - y0 + 1), cudaMemcpyDeviceToHost, ctx->stream));
    }
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    return VR_OK;
  }
  VR_CUDA(cudaMemcpyAsync(host_rgba, r->frame, (size_t)r->W * r->H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));  // also completes the bounding-box transfer of the k_primary before it
  if (own) r->host_epoch = (r->primary_valid && r->trace_mode >= 2 && r->row0 == 0 && r->row1 == r->H) ? r->primary_epoch : 0;
  return VR_OK;
}

extern "C" int vr_render_frame(vr_renderer* r, const float pos[3], const float dir[3], int32_t seed,
                               uint8_t* host_rgba) {
  VR_REQUIRE(r && pos && dir, "vr_render_frame: null argument");
  VR_TRY(volume_finish(r->vol));
  VR_TRY(check_flushed(r, "vr_render_frame"));
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_TRY(vrk_render(r, pos, dir, &seed, 1, true, true));
  return read_frame(r, host_rgba);
}

extern "C" int vr_render_frames(vr_renderer* r, const float pos[3], const float dir[3], const int32_t* seeds,
                                int n_frames, uint8_t* host_rgba) {
  VR_REQUIRE(r && pos && dir && seeds && n_frames > 0, "vr_render_frames: bad argument");
  VR_TRY(volume_finish(r->vol));
  VR_TRY(check_flushed(r, "vr_render_frames"));
  VR_CUDA(cudaSetDevice(r->ctx->device));
  // Only the last frame is observable, so the traces of a batch share one launch (their samples commute: integer
  // atomics) and the cache is resolved once at the end.  Which samples a voxel admits once it reaches the token cap
  // mid-batch is scheduling dependent — the same class of nondeterminism the reference has inside a single frame.
  for (int k = 0; k < n_frames; k += VR_MAX_BATCH) {
    const int nb = std::min(VR_MAX_BATCH, n_frames - k);
    VR_TRY(vrk_render(r, pos, dir, seeds + k, nb, true, k + nb == n_frames, k == 0));
  }
  return read_frame(r, host_rgba);
}

extern "C" int vr_renderer_resolve(vr_renderer* r, uint8_t* host_rgba) {
  VR_REQUIRE(r, "vr_renderer_resolve: null argument");
  VR_TRY(check_flushed(r, "vr_renderer_resolve"));
  VR_CUDA(cudaSetDevice(r->ctx->device));
  const float z[3] = {0, 0, 0};
  const int32_t zero = 0;
  VR_TRY(vrk_render(r, z, z, &zero, 1, false, true));
  return read_frame(r, host_rgba);
}

// ---- 2-D frame filter: 2d_image_filter.cl:6-43 -------------------------------------------------------------------------
static int filter2d_check(int kernel_size, float sigma, int mode) {
  VR_REQUIRE(mode == VR_FILTER2D_REFERENCE || mode == VR_FILTER2D_BILATERAL, "frame filter: unknown mode");
  VR_REQUIRE(kernel_size >= 0 && kernel_size <= (mode == VR_FILTER2D_REFERENCE ? 64 : 15),
             "frame filter: kernel_size out of range ([0,64] reference mode, [0,15] bilateral mode)");
  VR_REQUIRE(sigma > 0.0f && std::isfinite(sigma), "frame filter: sigma must be a positive finite number");
  return VR_OK;
}

extern "C" int vr_renderer_filter_frame(vr_renderer* r, int kernel_size, float sigma, int mode, uint8_t* host_rgba) {
  VR_REQUIRE(r, "vr_renderer_filter_frame: null argument");
  VR_TRY(filter2d_check(kernel_size, sigma, mode));
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)r->W * r->H * 4;
  // The filtered image goes to its own buffer: the traced frame must stay as it is, because with primary reuse across calls
  // (vr_renderer_set_primary_reuse(r, 2)) the environment pixels are written once per camera and only shaded pixels are
  // re-resolved — filtering in place would leave a blurred background in every later frame.
  if (!r->filtered) VR_CUDA(pool_alloc(ctx, &r->filtered, bytes));
  VR_TRY(vrk_filter2d(ctx, r->frame, r->filtered, r->W, r->H, kernel_size, sigma, mode));
  if (!host_rgba) return VR_OK;
  if (host_rgba == r->frame_host) r->host_epoch = 0;  // the host frame now holds the filtered image: the next pull is a full one
  VR_CUDA(cudaMemcpyAsync(host_rgba, r->filtered, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}
extern "C" void* vr_renderer_filtered_device_ptr(const vr_renderer* r) { return r ? (void*)r->filtered : nullptr; }

extern "C" int vr_image_filter(vr_ctx* ctx, const uint8_t* rgba_in, int w, int h, int kernel_size, float sigma, int mode,
                               uint8_t* rgba_out) {
  VR_REQUIRE(ctx && rgba_in && rgba_out, "vr_image_filter: null argument");
  VR_REQUIRE(w > 0 && h > 0 && (size_t)w * h < ((size_t)1 << 30), "vr_image_filter: bad image size");
  VR_TRY(filter2d_check(kernel_size, sigma, mode));
  VR_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)w * h * 4;
  uchar4 *src = nullptr, *dst = nullptr;
  int st = VR_OK;
  cudaError_t e = pool_alloc(ctx, &src, bytes);
  if (e == cudaSuccess) e = pool_alloc(ctx, &dst, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(src, rgba_in, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { vr_set_error("vr_image_filter: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  if (st == VR_OK) st = vrk_filter2d(ctx, src, dst, w, h, kernel_size, sigma, mode);
  if (st == VR_OK) {
    e = cudaMemcpyAsync(rgba_out, dst, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_image_filter: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, src);
  pool_free(ctx, dst);
  return st;
}

extern "C" uint8_t* vr_renderer_host_frame(vr_renderer* r) { return r ? r->frame_host : nullptr; }

extern "C" int vr_cache_download(const vr_renderer* r, uint16_t* out) {
  VR_REQUIRE(r && out && r->cache, "vr_cache_download: no cache");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_CUDA(cudaMemcpyAsync(out, r->cache, r->cache_voxels * 8, cudaMemcpyDeviceToHost, r->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(r->ctx->stream));
  return VR_OK;
}

// per pixel: voxel number of the primary hit the last trace found (utility.cl:21 order), 0xFFFFFFFF for environment pixels
extern "C" int vr_renderer_hit_download(const vr_renderer* r, uint32_t* out) {
  VR_REQUIRE(r && out && r->hit, "vr_renderer_hit_download: null argument");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_CUDA(cudaMemcpyAsync(out, r->hit, (size_t)r->W * r->H * 4, cudaMemcpyDeviceToHost, r->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(r->ctx->stream));
  return VR_OK;
}

// the cache entries of n given voxels (4 ushort each): parity checks at sizes where the whole cache is gigabytes
extern "C" int vr_cache_download_at(const vr_renderer* r, const uint32_t* voxels, size_t n, uint16_t* out) {
  VR_REQUIRE(r && voxels && out && r->cache, "vr_cache_download_at: no cache");
  for (size_t i = 0; i < n; ++i) VR_REQUIRE(voxels[i] < r->cache_voxels, "vr_cache_download_at: voxel out of range");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  if (!n) return VR_OK;
  uint32_t* idx = nullptr;
  uint2* ent = nullptr;
  VR_CUDA(pool_alloc(ctx, &idx, n * 4));
  VR_CUDA(pool_alloc(ctx, &ent, n * 8));
  int st = VR_OK;
  cudaError_t e = cudaMemcpyAsync(idx, voxels, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { vr_set_error("vr_cache_download_at: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  if (st == VR_OK) st = vrk_cache_gather(ctx, r->cache, idx, n, ent);
  if (st == VR_OK) {
    e = cudaMemcpyAsync(out, ent, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_cache_download_at: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, idx); pool_free(ctx, ent);
  return st;
}

extern "C" const vr_sdf* vr_renderer_sdf(const vr_renderer* r) { return r ? r->sdf : nullptr; }
extern "C" int vr_renderer_last_flush_kept_fields(const vr_renderer* r) { return r && r->fields_kept ? 1 : 0; }

extern "C" int vr_renderer_set_token_cap(vr_renderer* r, int cap) {
  VR_REQUIRE(r && cap >= 1 && cap <= 256, "vr_renderer_set_token_cap: cap must be in [1,256]");
  r->token_cap = cap;
  return VR_OK;
}

extern "C" int vr_renderer_set_trace_mode(vr_renderer* r, int mode) {
#ifdef VR_AB
  VR_REQUIRE(r && mode >= 0 && mode <= 3, "vr_renderer_set_trace_mode: mode must be 0, 1, 2 (or 3 in this A/B build)");
#else
  VR_REQUIRE(r && mode >= 0 && mode <= 2, "vr_renderer_set_trace_mode: mode must be 0, 1 or 2");
#endif
  r->trace_mode = mode;
  r->primary_valid = false;
  return VR_OK;
}

extern "C" int vr_renderer_set_primary_reuse(vr_renderer* r, int level) {
  VR_REQUIRE(r && (level == 1 || level == 2), "vr_renderer_set_primary_reuse: level must be 1 or 2");
  r->primary_across_calls = level == 2;
  r->primary_valid = false;
  return VR_OK;
}

extern "C" int vr_renderer_set_rows(vr_renderer* r, int y0, int y1) {
  VR_REQUIRE(r && y0 >= 0 && y1 <= r->H && y0 <= y1, "vr_renderer_set_rows: rows out of range");
  r->row0 = y0; r->row1 = y1;
  return VR_OK;
}

extern "C" void* vr_renderer_cache_device_ptr(const vr_renderer* r) {
  if (r) const_cast<vr_renderer*>(r)->cache_exposed = true;  // the caller may write anywhere: frame resets clear everything from now on
  return r ? (void*)r->cache : nullptr;
}
extern "C" size_t vr_renderer_cache_bytes(const vr_renderer* r) { return r ? r->cache_voxels * 8 : 0; }
extern "C" void* vr_renderer_frame_device_ptr(const vr_renderer* r) { return r ? (void*)r->frame : nullptr; }

extern "C" int vr_renderer_enable_counters(vr_renderer* r, int enable) {
  VR_REQUIRE(r, "vr_renderer_enable_counters: null argument");
  r->count = enable != 0;
  return VR_OK;
}

extern "C" int vr_renderer_counters(const vr_renderer* r, uint64_t out[6], int reset) {
  VR_REQUIRE(r && out, "vr_renderer_counters: null argument");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_CUDA(cudaMemcpyAsync(out, r->counters, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, r->ctx->stream));
  if (reset) VR_CUDA(cudaMemsetAsync(r->counters, 0, 6 * sizeof(uint64_t), r->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(r->ctx->stream));
  return VR_OK;
}

static int ensure_xchg(vr_renderer* r) {
  if (!r->xchg) VR_CUDA(pool_alloc(r->ctx, &r->xchg, (size_t)r->W * r->H * sizeof(uint4)));  // sized for the wide form of vr_cache_allreduce
  return VR_OK;
}
extern "C" int vr_renderer_xchg_gather(vr_renderer* r) {
  VR_REQUIRE(r, "vr_renderer_xchg_gather: null argument");
  VR_TRY(check_flushed(r, "vr_renderer_xchg_gather"));
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_TRY(ensure_xchg(r));
  return vrk_xchg(r, r->xchg, false);
}
extern "C" int vr_renderer_xchg_scatter(vr_renderer* r) {
  VR_REQUIRE(r && r->xchg, "vr_renderer_xchg_scatter: nothing gathered");
  VR_TRY(check_flushed(r, "vr_renderer_xchg_scatter"));
  VR_CUDA(cudaSetDevice(r->ctx->device));
  return vrk_xchg(r, r->xchg, true);
}
extern "C" void* vr_renderer_xchg_device_ptr(vr_renderer* r) {
  if (!r) return nullptr;
  cudaSetDevice(r->ctx->device);
  if (ensure_xchg(r) != VR_OK) return nullptr;
  return r->xchg;
}
extern "C" size_t vr_renderer_xchg_bytes(const vr_renderer* r) { return r ? (size_t)r->W * r->H * sizeof(uint2) : 0; }

extern "C" int vr_renderer_enable_timing(vr_renderer* r, int enable) {
  VR_REQUIRE(r, "vr_renderer_enable_timing: null argument");
  r->timing = enable != 0;
  return VR_OK;
}

extern "C" int vr_renderer_kernel_times(vr_renderer* r, double out_ms[2], int* n_frames, int reset) {
  VR_REQUIRE(r && out_ms && n_frames, "vr_renderer_kernel_times: null argument");
  VR_CUDA(cudaSetDevice(r->ctx->device));
  VR_CUDA(cudaStreamSynchronize(r->ctx->stream));
  out_ms[0] = out_ms[1] = 0.0;
  for (size_t i = 0; i + 3 <= r->ev_used; i += 3) {
    float a = 0.f, b = 0.f;
    VR_CUDA(cudaEventElapsedTime(&a, r->ev[i], r->ev[i + 1]));
    VR_CUDA(cudaEventElapsedTime(&b, r->ev[i + 1], r->ev[i + 2]));
    out_ms[0] += a;
    out_ms[1] += b;
  }
  *n_frames = 0;
  for (int f : r->ev_frames) *n_frames += f;
  if (reset) { r->ev_used = 0; r->ev_frames.clear(); }
  return VR_OK;
}

// ---- render_tf: renderer.cpp:45-124 ---------------------------------------------------------------------------------
extern "C" int vr_volume_set_value_clip(vr_volume* v, int lo, int hi) {
  VR_REQUIRE(v, "vr_volume_set_value_clip: null argument");
  v->value_clip[0] = lo; v->value_clip[1] = hi;
  return VR_OK;
}
extern "C" int vr_volume_set_gradient_clip(vr_volume* v, int lo, int hi) {
  VR_REQUIRE(v, "vr_volume_set_gradient_clip: null argument");
  v->gradient_clip[0] = lo; v->gradient_clip[1] = hi;
  return VR_OK;
}
// get_volume_stats(), reference_volume.cpp:82-88,110-112
extern "C" int vr_volume_clipped_stats(const vr_volume* v, float out[4]) {
  VR_REQUIRE(v && out, "vr_volume_clipped_stats: null argument");
  VR_TRY(volume_finish(v));
  out[0] = (float)std::max(v->value_clip[0], v->stats[0]);
  out[1] = (float)std::min(v->value_clip[1], v->stats[1]);
  out[2] = (float)std::max(v->gradient_clip[0], v->stats[2]);
  out[3] = (float)std::min(v->gradient_clip[1], v->stats[3]);
  return VR_OK;
}

extern "C" int vr_render_tf(vr_renderer* r, int width, int height, uint8_t* rgba_out) {
  VR_REQUIRE(r && r->vol && rgba_out, "vr_render_tf: no scene bound");
  VR_TRY(volume_finish(r->vol));
  VR_REQUIRE(width > 0 && height > 0 && (size_t)width * height < ((size_t)1 << 31), "vr_render_tf: bad size");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  const size_t nb = (size_t)width * height;
  float range[4];
  VR_TRY(vr_volume_clipped_stats(r->vol, range));
  uint32_t* bins = nullptr;
  int* scratch = nullptr;
  uchar4* img = nullptr;
  int status = VR_OK;
  cudaError_t e = pool_alloc(ctx, &bins, nb * 4);
  if (e == cudaSuccess) e = pool_alloc(ctx, &img, nb * 4);
  if (e == cudaSuccess) e = pool_alloc(ctx, &scratch, 2049 * sizeof(int));
  if (e != cudaSuccess) { vr_set_error("vr_render_tf: %s", cudaGetErrorString(e)); status = VR_ERR_CUDA; }
  if (status == VR_OK)
    status = r->vol->sampling == VR_SAMPLING_HW_LINEAR
                 ? vrk_histogram_linear(ctx, r->vol->box, r->vol->box_px, r->vol->current(), r->vol->nx, r->vol->ny, r->vol->nz, width, height,
                                        range, bins, r->vol->zlo, r->vol->zhi, r->vol->stats[0], r->vol->raw_range[0], r->vol->raw_range[1])
                 : vrk_histogram(ctx, r->vol->current(), r->vol->nx, r->vol->ny, r->vol->nz, width, height, range, bins, r->vol->zlo,
                                 r->vol->zhi, r->vol->stats[0], r->vol->stats[1]);
  // renderer.cpp:65-96 without the host round trip: rounding, distinct-value ranking and colouring stay on the device
  if (status == VR_OK) status = vrk_tf_image(ctx, reinterpret_cast<int32_t*>(bins), scratch, width, height, img);
  if (status == VR_OK) {
    e = cudaMemcpyAsync(rgba_out, img, nb * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_render_tf: %s", cudaGetErrorString(e)); status = VR_ERR_CUDA; }
  }
  pool_free(ctx, scratch);
  pool_free(ctx, bins);
  pool_free(ctx, img);
  return status;
}

// ---- schedule tuning, hw-linear diagnostics, device RNG known answers ---------------------------------------------------------
extern "C" int vr_renderer_set_tuning(vr_renderer* r, const char* key, int value) {
  VR_REQUIRE(r && key, "vr_renderer_set_tuning: null argument");
  auto& t = r->tune;
  if (!strcmp(key, "pixel_major")) { VR_REQUIRE(value >= 0 && value <= 4096, "pixel_major out of range"); t.pixel_major = value; }
  else if (!strcmp(key, "rule_a")) { VR_REQUIRE(value >= 1 && value <= 1024, "rule_a out of range"); t.rule[0] = value; }
  else if (!strcmp(key, "rule_b")) { VR_REQUIRE(value >= 1 && value <= 1024, "rule_b out of range"); t.rule[1] = value; }
  else if (!strcmp(key, "sm_k")) { VR_REQUIRE(value >= 1 && value <= 70, "sm_k out of range"); t.sm_k = value; }
  else if (!strcmp(key, "sm_leave")) { VR_REQUIRE(value >= 0 && value <= 32, "sm_leave out of range"); t.sm_leave = value; }
  else if (!strcmp(key, "surf")) { VR_REQUIRE(value == 0 || value == 1, "surf must be 0 or 1"); t.surf = value; }
  else if (!strcmp(key, "lin_sched")) { VR_REQUIRE(value == 0 || value == 1, "lin_sched must be 0 or 1"); t.lin_sched = value; }
  else if (!strcmp(key, "lin_w_fast")) { VR_REQUIRE(value >= 1 && value <= 1024, "lin_w_fast out of range"); t.lin_w[0] = value; }
  else if (!strcmp(key, "lin_w_slow")) { VR_REQUIRE(value >= 1 && value <= 1024, "lin_w_slow out of range"); t.lin_w[1] = value; }
  else if (!strcmp(key, "lin_w_event")) { VR_REQUIRE(value >= 1 && value <= 1024, "lin_w_event out of range"); t.lin_w[2] = value; }
  else if (!strcmp(key, "steps_per_check")) { VR_REQUIRE(value >= 1 && value <= 16, "steps_per_check out of range"); t.spc = value; }
  else if (!strcmp(key, "pt_ctas") || !strcmp(key, "pt_slots") || !strcmp(key, "pt2_ctas")) {
#ifdef VR_AB
    if (!strcmp(key, "pt_slots")) { VR_REQUIRE(value == 1 || value == 2, "pt_slots must be 1 or 2"); t.pt_slots = value; }
    else if (!strcmp(key, "pt2_ctas")) { VR_REQUIRE(value == 6 || value == 7 || value == 8 || value == 10, "pt2_ctas must be 6, 7, 8 or 10"); t.pt2_ctas = value; }
    else t.pt_ctas = value;
#else
    vr_set_error("vr_renderer_set_tuning: %s needs the A/B build of the library (make ab)", key);
    return VR_ERR_INVALID;
#endif
  } else { vr_set_error("vr_renderer_set_tuning: unknown key '%s'", key); return VR_ERR_INVALID; }
  return VR_OK;
}

extern "C" int vr_renderer_quiet_download(const vr_renderer* r, uint8_t* out) {
  VR_REQUIRE(r && out, "vr_renderer_quiet_download: null argument");
  VR_REQUIRE(r->lin_surf && r->flushed_vol, "vr_renderer_quiet_download: no step field (hw-linear sampling + flush)");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  const vr_volume* v = r->flushed_vol;
  uint8_t* dev = nullptr;
  VR_CUDA(pool_alloc(ctx, &dev, v->count()));
  int st = vrk_lin_field_masks(ctx, r->lin_surf, v->nx, v->ny, v->nz, dev);
  if (st == VR_OK) {
    cudaError_t e = cudaMemcpyAsync(out, dev, v->count(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_renderer_quiet_download: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, dev);
  return st;
}

extern "C" int vr_debug_linear_fetch(const vr_renderer* r, const float* xyz, int n, int32_t* out) {
  VR_REQUIRE(r && xyz && out && n > 0, "vr_debug_linear_fetch: bad argument");
  VR_REQUIRE(r->vol_tex, "vr_debug_linear_fetch: no volume texture (hw-linear sampling + flush)");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  float* d_xyz = nullptr;
  int32_t* d_out = nullptr;
  VR_CUDA(pool_alloc(ctx, &d_xyz, (size_t)n * 12));
  VR_CUDA(pool_alloc(ctx, &d_out, (size_t)n * 4));
  int st = VR_OK;
  cudaError_t e = cudaMemcpyAsync(d_xyz, xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) { vr_set_error("vr_debug_linear_fetch: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  if (st == VR_OK) st = vrk_linear_fetch(ctx, r->vol_tex, d_xyz, n, d_out);
  if (st == VR_OK) {
    e = cudaMemcpyAsync(out, d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_debug_linear_fetch: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, d_xyz); pool_free(ctx, d_out);
  return st;
}

extern "C" int vr_debug_rng_dump(vr_ctx* ctx, const int32_t* seeds, const uint32_t* gid_xy, const float* normal_rough, int n,
                                 int32_t* ra_out, int32_t* comp_out, float* dir_out) {
  VR_REQUIRE(ctx && seeds && gid_xy && normal_rough && ra_out && comp_out && dir_out && n > 0, "vr_debug_rng_dump: bad argument");
  VR_CUDA(cudaSetDevice(ctx->device));
  int32_t *d_seeds = nullptr, *d_ra = nullptr, *d_comp = nullptr;
  uint32_t* d_gid = nullptr;
  float *d_nr = nullptr, *d_dir = nullptr;
  const size_t N = (size_t)n;
  cudaError_t e = pool_alloc(ctx, &d_seeds, 4 * N);
  if (e == cudaSuccess) e = pool_alloc(ctx, &d_gid, 8 * N);
  if (e == cudaSuccess) e = pool_alloc(ctx, &d_nr, 16 * N);
  if (e == cudaSuccess) e = pool_alloc(ctx, &d_ra, 12 * N);
  if (e == cudaSuccess) e = pool_alloc(ctx, &d_comp, 12 * N);
  if (e == cudaSuccess) e = pool_alloc(ctx, &d_dir, 12 * N);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_seeds, seeds, 4 * N, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_gid, gid_xy, 8 * N, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_nr, normal_rough, 16 * N, cudaMemcpyHostToDevice, ctx->stream);
  int st = VR_OK;
  if (e != cudaSuccess) { vr_set_error("vr_debug_rng_dump: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  if (st == VR_OK) st = vrk_rng_dump(ctx, d_seeds, d_gid, n, d_nr, d_ra, d_comp, d_dir);
  if (st == VR_OK) {
    e = cudaMemcpyAsync(ra_out, d_ra, 12 * N, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(comp_out, d_comp, 12 * N, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dir_out, d_dir, 12 * N, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_debug_rng_dump: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  pool_free(ctx, d_seeds); pool_free(ctx, d_gid); pool_free(ctx, d_nr); pool_free(ctx, d_ra); pool_free(ctx, d_comp); pool_free(ctx, d_dir);
  return st;
}
