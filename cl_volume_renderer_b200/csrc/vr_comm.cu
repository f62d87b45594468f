// vr_comm.cu — the multi-GPU entry points of include/vr.h: one process (or thread) per GPU, NCCL over NVLink / NVSwitch,
// every collective enqueued on the context's stream by this library, so a C++ host needs nothing but the C-ABI.
// The reference is single-device; SURVEY.md 8(b)/(e) names what is exported here.
//
//   vr_comm_init               communicator of the ranks that hold one vr_ctx each
//   vr_cache_allreduce         spp split: sum of the touched voxel-cache entries + resolve (compact, one entry per shaded pixel)
//   vr_frame_allgather         image-tile split: every rank's row blocks -> the whole frame on every rank
//   vr_volume_upload_sharded   ingest: every rank uploads 1/N of the planes over PCIe, the rest arrives over NVLink
//   vr_sdf_build_sharded       z-slab SDF build with halo swaps of the bit volume (ncclSend/ncclRecv) + gather of the field
//   vr_histogram_sharded / vr_volume_filter_sharded   z-slab partials + all-reduce / gather
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include "vr_sdf_common.cuh"

// NCCL is bound at run time, on the first vr_comm_* call, not at link time: a single-GPU host needs no libnccl at all, and a
// process that already carries an NCCL (a Python host that imported torch, which ships its own libnccl.so.2) keeps using that
// one copy — dlopen by SONAME returns the library that is already loaded instead of a second, older one.
namespace {
struct NcclApi {
  void* lib = nullptr;
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllReduce) AllReduce = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclBroadcast) Broadcast = nullptr;
  decltype(&::ncclSend) Send = nullptr;
  decltype(&::ncclRecv) Recv = nullptr;
  decltype(&::ncclGroupStart) GroupStart = nullptr;
  decltype(&::ncclGroupEnd) GroupEnd = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
  decltype(&::ncclCommSplit) CommSplit = nullptr;  // optional (NCCL >= 2.18): the copy-stream communicator
  bool ok = false;
} g_nccl;

int nccl_bind() {
  if (g_nccl.ok) return VR_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { vr_set_error("vr_comm: libnccl.so.2 not found (%s)", dlerror()); return VR_ERR_INVALID; }
  g_nccl.lib = h;
  bool all = true;
#define VR_BIND(name)                                                                \
  g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name));     \
  all = all && g_nccl.name != nullptr
  VR_BIND(GetUniqueId); VR_BIND(CommInitRank); VR_BIND(CommDestroy); VR_BIND(AllReduce); VR_BIND(AllGather); VR_BIND(Broadcast);
  VR_BIND(Send); VR_BIND(Recv); VR_BIND(GroupStart); VR_BIND(GroupEnd); VR_BIND(GetErrorString);
#undef VR_BIND
  g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(h, "ncclCommSplit"));
  if (!all) { vr_set_error("vr_comm: libnccl.so.2 lacks an expected symbol"); return VR_ERR_INVALID; }
  g_nccl.ok = true;
  return VR_OK;
}
}  // namespace
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclBroadcast g_nccl.Broadcast
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString

#define VR_NCCL(call)                                                                               \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != ncclSuccess) {                                                                       \
      vr_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));     \
      return VR_ERR_CUDA;                                                                           \
    }                                                                                               \
  } while (0)

static inline ncclComm_t comm_of(const vr_ctx* c) { return (ncclComm_t)c->comm; }

extern "C" int vr_comm_unique_id(uint8_t id[VR_COMM_ID_BYTES]) {
  VR_REQUIRE(id, "vr_comm_unique_id: null argument");
  static_assert(VR_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
  VR_TRY(nccl_bind());
  ncclUniqueId u;
  VR_NCCL(ncclGetUniqueId(&u));
  memcpy(id, u.internal, VR_COMM_ID_BYTES);
  return VR_OK;
}

extern "C" int vr_comm_init(vr_ctx* ctx, int rank, int nranks, const uint8_t id[VR_COMM_ID_BYTES]) {
  VR_REQUIRE(ctx && id && nranks >= 1 && rank >= 0 && rank < nranks, "vr_comm_init: bad argument");
  VR_REQUIRE(!ctx->comm, "vr_comm_init: the context already has a communicator");
  VR_TRY(nccl_bind());
  VR_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(u.internal, id, VR_COMM_ID_BYTES);
  ncclComm_t c = nullptr;
  VR_NCCL(ncclCommInitRank(&c, nranks, u, rank));
  ctx->comm = c; ctx->comm_rank = rank; ctx->comm_size = nranks;
  // collectives of the asynchronous sharded ingest run on the copy stream beside the compute stream's: they need a communicator
  // of their own (operations of ONE communicator must be issued in one order on all ranks)
  if (nranks > 1 && g_nccl.CommSplit) {
    ncclComm_t c2 = nullptr;
    if (g_nccl.CommSplit(c, 0, rank, &c2, nullptr) == ncclSuccess) ctx->comm_copy = c2;
  }
  return VR_OK;
}

void vr_comm_release(vr_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  if (ctx->comm_copy) ncclCommDestroy((ncclComm_t)ctx->comm_copy);
  ctx->comm_copy = nullptr;
  ncclCommDestroy(comm_of(ctx));
  ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_size = 1;
}
extern "C" void vr_comm_destroy(vr_ctx* ctx) { vr_comm_release(ctx); }
extern "C" int vr_comm_rank(const vr_ctx* ctx) { return ctx ? ctx->comm_rank : 0; }
extern "C" int vr_comm_size(const vr_ctx* ctx) { return ctx ? ctx->comm_size : 1; }

// host values in, reduced host values out (timings, checksums, stats): dtype 0 int32, 1 uint32, 2 float64; op 0 sum, 1 min, 2 max
extern "C" int vr_comm_allreduce_host(vr_ctx* ctx, void* values, int count, int dtype, int op) {
  VR_REQUIRE(ctx && values && count > 0 && dtype >= 0 && dtype <= 2 && op >= 0 && op <= 2, "vr_comm_allreduce_host: bad argument");
  const size_t bytes = (size_t)count * (dtype == 2 ? 8 : 4);
  VR_REQUIRE(bytes <= 2048, "vr_comm_allreduce_host: at most 2 KiB");
  if (!ctx->comm || ctx->comm_size == 1) return VR_OK;
  VR_CUDA(cudaSetDevice(ctx->device));
  char* dev = reinterpret_cast<char*>(ctx->scratch) + 2048;  // upper half of the context's 4 KiB scratch
  char* pin = reinterpret_cast<char*>(ctx->scratch_host) + 2048;
  memcpy(pin, values, bytes);
  VR_CUDA(cudaMemcpyAsync(dev, pin, bytes, cudaMemcpyHostToDevice, ctx->stream));
  const ncclDataType_t t = dtype == 0 ? ncclInt32 : (dtype == 1 ? ncclUint32 : ncclFloat64);
  const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
  VR_NCCL(ncclAllReduce(dev, dev, (size_t)count, t, o, comm_of(ctx), ctx->stream));
  VR_CUDA(cudaMemcpyAsync(pin, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(values, pin, bytes);
  return VR_OK;
}

extern "C" int vr_comm_barrier(vr_ctx* ctx) {
  int32_t one = 1;
  return vr_comm_allreduce_host(ctx, &one, 1, 0, 0);
}

// z-partition used by every sharded call: brick-aligned (multiples of 8 planes), so a rank's part of the bricked SDF field is a
// contiguous range of brick layers; the last rank takes the remainder
extern "C" int vr_comm_slab(const vr_ctx* ctx, int nz, int rank, int* z0, int* z1) {
  VR_REQUIRE(ctx && z0 && z1 && nz > 0, "vr_comm_slab: bad argument");
  const int n = ctx->comm_size;
  VR_REQUIRE(rank >= 0 && rank < n, "vr_comm_slab: rank out of range");
  const int layers = (nz + 7) / 8;
  const int base = layers / n, rem = layers % n;
  const int l0 = rank * base + std::min(rank, rem), l1 = l0 + base + (rank < rem ? 1 : 0);
  *z0 = std::min(nz, l0 * 8);
  *z1 = std::min(nz, l1 * 8);
  return VR_OK;
}

// ---- spp split: vr_cache_allreduce ------------------------------------------------------------------------------------------
extern "C" int vr_cache_allreduce(vr_renderer* r, uint8_t* host_rgba) {
  VR_REQUIRE(r && r->sdf && r->cache, "vr_cache_allreduce: call vr_renderer_flush first");
  VR_REQUIRE(r->row0 == 0 && r->row1 == r->H && r->blk_n <= 1, "vr_cache_allreduce: the spp split traces the whole frame on every rank");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  // packed 16-bit lanes cannot carry as long as the caps of all ranks add up to at most 256 tokens per voxel; otherwise the
  // lanes travel as uint32 words and the frame is resolved from the wide sums
  const bool wide = (long long)r->token_cap * ctx->comm_size > 256;
  unsigned* count_dev = nullptr;
  VR_TRY(vrk_xc_gather(r, &count_dev, wide));
  if (ctx->comm && ctx->comm_size > 1) {
    // the NCCL count = shaded pixels, identical on all ranks (same camera, same hit buffer).  The hit buffer is a function of
    // camera, rows and flushed scene, so the count is read back (4 bytes + a stream synchronisation) once per such signature
    // and a progressive loop enqueues its exchanges without ever waiting for the device.
    const bool same = r->xc_sig_valid && r->xc_flush == r->flush_count && !memcmp(r->xc_pos, r->dirty_pos, 12) &&
                      !memcmp(r->xc_dir, r->dirty_dir, 12) && r->cache_dirty == 1;
    if (!same) {
      unsigned* pin = reinterpret_cast<unsigned*>(ctx->scratch_host);
      VR_CUDA(cudaMemcpyAsync(pin, count_dev, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      r->xc_count_host = pin[0];
      r->xc_sig_valid = r->cache_dirty == 1;  // one camera since the last reset: dirty_pos / dirty_dir describe the hit buffer
      r->xc_flush = r->flush_count;
      memcpy(r->xc_pos, r->dirty_pos, 12); memcpy(r->xc_dir, r->dirty_dir, 12);
    }
    const size_t words = (size_t)r->xc_count_host * (wide ? 4 : 2);
    if (words) VR_NCCL(ncclAllReduce(r->xchg, r->xchg, words, ncclUint32, ncclSum, comm_of(ctx), ctx->stream));
  }
  VR_TRY(vrk_xc_scatter_resolve(r, wide));
  if (!host_rgba) return VR_OK;
  if (host_rgba == r->frame_host) r->host_epoch = 0;  // full pull outside read_frame: the next incremental pull starts over
  VR_CUDA(cudaMemcpyAsync(host_rgba, r->frame, (size_t)r->W * r->H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

// ---- image-tile split: vr_renderer_set_row_blocks + vr_frame_allgather ------------------------------------------------------------
extern "C" int vr_renderer_set_row_blocks(vr_renderer* r, int block_rows, int rank, int nranks) {
  VR_REQUIRE(r && nranks >= 1 && rank >= 0 && rank < nranks && (nranks == 1 || block_rows >= 1), "vr_renderer_set_row_blocks: bad argument");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  r->blk_rows = nranks > 1 ? block_rows : 0; r->blk_rank = nranks > 1 ? rank : 0; r->blk_n = nranks;
  r->primary_valid = false;
  if (r->cache_dirty) r->cache_dirty = 2;  // the sparse reset walks the hit buffer, which is cleared here
  const size_t px = (size_t)r->W * r->H;
  VR_CUDA(cudaMemsetAsync(r->hit, 0xFF, px * 4, ctx->stream));
  VR_CUDA(cudaMemsetAsync(r->frame, 0, px * 4, ctx->stream));
  return VR_OK;
}

// rank's blocks, in order, into a dense chunk of `per` blocks (missing rows at the bottom edge stay zero)
__global__ void __launch_bounds__(256) k_fa_pack(const uint32_t* __restrict__ frame, int W, int H, int block_rows, int rank, int n,
                                                 int per, uint32_t* __restrict__ out) {
  const size_t total = (size_t)per * block_rows * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const size_t t = i / W;
    const int rr = (int)(t % block_rows), j = (int)(t / block_rows);
    const int y = (j * n + rank) * block_rows + rr;
    out[i] = y < H ? frame[(size_t)y * W + x] : 0u;
  }
}
__global__ void __launch_bounds__(256) k_fa_unpack(const uint32_t* __restrict__ gathered, int W, int H, int block_rows, int n, int per,
                                                   uint32_t* __restrict__ frame) {
  const size_t total = (size_t)W * H, chunk = (size_t)per * block_rows * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)(i / W);
    const int b = y / block_rows, owner = b % n, j = b / n;
    frame[i] = gathered[(size_t)owner * chunk + ((size_t)j * block_rows + (y - b * block_rows)) * W + x];
  }
}

extern "C" int vr_frame_allgather(vr_renderer* r, uint8_t* host_rgba) {
  VR_REQUIRE(r, "vr_frame_allgather: null argument");
  vr_ctx* ctx = r->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  const int n = ctx->comm_size;
  if (ctx->comm && n > 1) {
    VR_REQUIRE(r->blk_n == n && r->blk_rank == ctx->comm_rank && r->blk_rows >= 1,
               "vr_frame_allgather: call vr_renderer_set_row_blocks(r, rows, vr_comm_rank, vr_comm_size) first");
    const int nblocks = (r->H + r->blk_rows - 1) / r->blk_rows, per = (nblocks + n - 1) / n;
    const size_t chunk = (size_t)per * r->blk_rows * r->W;  // pixels
    const size_t need = chunk * 4 * ((size_t)n + 1);
    if (r->gather_bytes < need) {
      if (r->gather_buf) VR_CUDA(cudaFreeAsync(r->gather_buf, ctx->stream));
      r->gather_buf = nullptr; r->gather_bytes = 0;
      VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&r->gather_buf), need, ctx->stream));
      r->gather_bytes = need;
    }
    uint32_t* mine = r->gather_buf;
    uint32_t* all = r->gather_buf + chunk;
    const unsigned blocks = (unsigned)std::min<size_t>(div_up(chunk, 256), (size_t)ctx->sm_count * 8);
    k_fa_pack<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(r->frame), r->W, r->H, r->blk_rows, r->blk_rank, n, per, mine);
    VR_NCCL(ncclAllGather(mine, all, chunk, ncclUint32, comm_of(ctx), ctx->stream));
    const unsigned ub = (unsigned)std::min<size_t>(div_up((size_t)r->W * r->H, 256), (size_t)ctx->sm_count * 8);
    k_fa_unpack<<<ub, 256, 0, ctx->stream>>>(all, r->W, r->H, r->blk_rows, n, per, reinterpret_cast<uint32_t*>(r->frame));
    ctx->launches += 2;
    VR_CUDA(cudaGetLastError());
    // the frame now holds other ranks' rows: the next trace from this camera must rewrite its environment pixels
    r->primary_valid = false;
  }
  if (!host_rgba) return VR_OK;
  if (host_rgba == r->frame_host) r->host_epoch = 0;  // full pull outside read_frame: the next incremental pull starts over
  VR_CUDA(cudaMemcpyAsync(host_rgba, r->frame, (size_t)r->W * r->H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  return VR_OK;
}

// ---- sharded ingest -----------------------------------------------------------------------------------------------------------
// in-place gather of per-rank byte ranges of one buffer: a group of broadcasts, one per rank (ranges may differ in size)
static int gather_ranges(vr_ctx* ctx, void* base, const size_t* off, const size_t* len, bool on_copy_stream = false) {
  ncclComm_t comm = on_copy_stream ? (ncclComm_t)ctx->comm_copy : comm_of(ctx);
  cudaStream_t stream = on_copy_stream ? ctx->copy_stream : ctx->stream;
  VR_NCCL(ncclGroupStart());
  for (int k = 0; k < ctx->comm_size; ++k)
    if (len[k]) {
      char* p = reinterpret_cast<char*>(base) + off[k];
      ncclResult_t e = ncclBroadcast(p, p, len[k], ncclUint8, k, comm, stream);
      if (e != ncclSuccess) { ncclGroupEnd(); vr_set_error("ncclBroadcast: %s", ncclGetErrorString(e)); return VR_ERR_CUDA; }
    }
  VR_NCCL(ncclGroupEnd());
  return VR_OK;
}

extern "C" int vr_volume_upload_sharded(vr_ctx* ctx, const int16_t* own_planes, int nx, int ny, int nz, vr_volume** out) {
  VR_REQUIRE(ctx && own_planes && out, "vr_volume_upload_sharded: null argument");
  VR_REQUIRE(nx > 0 && ny > 0 && nz > 0, "vr_volume_upload_sharded: dimensions must be positive");
  VR_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 32) - 1, "vr_volume_upload_sharded: more than 2^32-2 voxels");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_volume* v = new (std::nothrow) vr_volume();
  if (!v) return VR_ERR_NOMEM;
  v->ctx = ctx;
  v->onx = v->nx = nx; v->ony = v->ny = ny; v->onz = v->nz = nz;
  v->zlo = 0; v->zhi = nz;
  const size_t plane = (size_t)nx * ny * sizeof(int16_t);
  int st = VR_OK;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&v->original), plane * nz, ctx->stream);
  if (e != cudaSuccess) { vr_set_error("vr_volume_upload_sharded: %s", cudaGetErrorString(e)); delete v; return VR_ERR_CUDA; }
  size_t off[64], len[64];
  const int n = ctx->comm_size;
  if (n > 64) { st = VR_ERR_INVALID; vr_set_error("vr_volume_upload_sharded: more than 64 ranks"); }
  for (int k = 0; k < n && st == VR_OK; ++k) {
    int z0, z1;
    st = vr_comm_slab(ctx, nz, k, &z0, &z1);
    off[k] = plane * z0; len[k] = plane * (z1 - z0);
  }
  if (st == VR_OK && len[ctx->comm_rank]) {
    e = cudaMemcpyAsync(reinterpret_cast<char*>(v->original) + off[ctx->comm_rank], own_planes, len[ctx->comm_rank], cudaMemcpyHostToDevice,
                        ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_volume_upload_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  if (st == VR_OK && ctx->comm && n > 1) st = gather_ranges(ctx, v->original, off, len);
  // fetch_stats (reference_volume.cpp:22-41): every rank reduces its own planes (their gradient taps reach the gathered
  // neighbours), one MIN all-reduce of {min v, -max v, min g, -max g} combines them
  if (st == VR_OK) {
    int z0 = 0, z1 = nz;
    if (ctx->comm && n > 1) st = vr_comm_slab(ctx, nz, ctx->comm_rank, &z0, &z1);
    if (st == VR_OK && z1 > z0) st = vrk_fetch_stats(ctx, v->original, nx, ny, nz, v->stats, z0, z1);
    else if (st == VR_OK) { v->stats[0] = v->stats[2] = INT32_MAX; v->stats[1] = v->stats[3] = INT32_MIN; }
    if (st == VR_OK && ctx->comm && n > 1) {
      int32_t t[4] = {v->stats[0], v->stats[1] == INT32_MIN ? INT32_MAX : -v->stats[1], v->stats[2], v->stats[3] == INT32_MIN ? INT32_MAX : -v->stats[3]};
      st = vr_comm_allreduce_host(ctx, t, 4, 0, 1);
      v->stats[0] = t[0]; v->stats[1] = -t[1]; v->stats[2] = t[2]; v->stats[3] = -t[3];
    }
  }
  if (st != VR_OK) {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(v->original, ctx->stream);
    delete v;
    return st;
  }
  *out = v;
  return VR_OK;
}

// The same ingest without blocking: copy, gather and fetch_stats run on the context's copy stream (the gather on the copy-stream
// communicator) beside whatever the compute stream is doing — job k+1 arrives while job k builds its SDF and renders.  Like
// vr_volume_upload_async, every call that uses the volume waits for it; `own_planes` must stay valid until then.
extern "C" int vr_volume_upload_sharded_async(vr_ctx* ctx, const int16_t* own_planes, int nx, int ny, int nz, vr_volume** out) {
  VR_REQUIRE(ctx && own_planes && out, "vr_volume_upload_sharded_async: null argument");
  if (!ctx->comm || ctx->comm_size == 1) return vr_volume_upload_async(ctx, own_planes, nx, ny, nz, out);
  if (!ctx->comm_copy) return vr_volume_upload_sharded(ctx, own_planes, nx, ny, nz, out);  // NCCL without ncclCommSplit: blocking form
  VR_REQUIRE(nx > 0 && ny > 0 && nz > 0, "vr_volume_upload_sharded_async: dimensions must be positive");
  VR_REQUIRE((size_t)nx * ny * nz < ((size_t)1 << 32) - 1, "vr_volume_upload_sharded_async: more than 2^32-2 voxels");
  const int n = ctx->comm_size;
  VR_REQUIRE(n <= 64, "vr_volume_upload_sharded_async: more than 64 ranks");
  VR_CUDA(cudaSetDevice(ctx->device));
  vr_volume* v = new (std::nothrow) vr_volume();
  if (!v) return VR_ERR_NOMEM;
  v->ctx = ctx;
  v->onx = v->nx = nx; v->ony = v->ny = ny; v->onz = v->nz = nz;
  v->zlo = 0; v->zhi = nz;
  const size_t plane = (size_t)nx * ny * sizeof(int16_t);
  size_t off[64], len[64];
  int st = VR_OK;
  for (int k = 0; k < n && st == VR_OK; ++k) {
    int z0, z1;
    st = vr_comm_slab(ctx, nz, k, &z0, &z1);
    off[k] = plane * z0; len[k] = plane * (z1 - z0);
  }
  cudaError_t e = cudaSuccess;
  if (st == VR_OK) e = cudaMallocAsync(reinterpret_cast<void**>(&v->original), plane * nz, ctx->copy_stream);
  if (st == VR_OK && e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&v->stats_dev), 4 * sizeof(int32_t), ctx->copy_stream);
  if (st == VR_OK && e == cudaSuccess) e = pinned_acquire(ctx, reinterpret_cast<void**>(&v->stats_pin), 64);
  if (st == VR_OK && e == cudaSuccess) e = cudaEventCreateWithFlags(&v->ready, cudaEventDisableTiming);
  if (st == VR_OK && e == cudaSuccess && len[ctx->comm_rank])
    e = cudaMemcpyAsync(reinterpret_cast<char*>(v->original) + off[ctx->comm_rank], own_planes, len[ctx->comm_rank], cudaMemcpyHostToDevice,
                        ctx->copy_stream);
  if (st == VR_OK && e != cudaSuccess) { vr_set_error("vr_volume_upload_sharded_async: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  if (st == VR_OK) st = gather_ranges(ctx, v->original, off, len, true);
  if (st == VR_OK) st = vrk_fetch_stats_enqueue(ctx, ctx->copy_stream, v->original, nx, ny, nz, 0, nz, v->stats_dev, v->stats_pin);
  if (st == VR_OK && (e = cudaEventRecord(v->ready, ctx->copy_stream)) != cudaSuccess) {
    vr_set_error("vr_volume_upload_sharded_async: %s", cudaGetErrorString(e));
    st = VR_ERR_CUDA;
  }
  if (st != VR_OK) {
    cudaStreamSynchronize(ctx->copy_stream);
    if (v->original) cudaFreeAsync(v->original, ctx->copy_stream);
    if (v->stats_dev) cudaFreeAsync(v->stats_dev, ctx->copy_stream);
    if (v->stats_pin) pinned_release(ctx, v->stats_pin);
    if (v->ready) cudaEventDestroy(v->ready);
    delete v;
    return st;
  }
  v->pending = true;
  *out = v;
  return VR_OK;
}

// ---- z-slab SDF build ------------------------------------------------------------------------------------------------------
// The wave's results go stale from a slab's artificial ends inwards by one plane per level (plus two at the start: gradient taps
// of the event test, the band test), so a rank runs K levels, swaps the K + 2 boundary planes of the CURRENT bit volume with
// both z-neighbours (ncclSend / ncclRecv in one group, stream-ordered: no host synchronisation inside the build) and continues.
// Slabs are brick-aligned, the halo is 16 planes = two brick layers: a rank assembles its slab's bricked field and copies its
// own brick layers into place in the global field; one group of broadcasts completes the field on every rank.
#define VR_SDF_K 14
#define VR_SDF_HALO (VR_SDF_K + 2)

int vrk_sdf_build_sharded(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field, int* levels_out,
                          int* max_it_out, cudaSurfaceObject_t surf, bool gather) {
  const int n = ctx->comm_size, rank = ctx->comm_rank;
  const int max_it = std::min(std::max(nx, std::max(ny, nz)) / 2, 127);  // signed_distance_field.cpp:11 on the GLOBAL volume
  int z0s[64], z1s[64];
  bool ok = ctx->comm && n > 1 && n <= 64;
  for (int k = 0; k < n && ok; ++k) {
    if (vr_comm_slab(ctx, nz, k, &z0s[k], &z1s[k]) != VR_OK) ok = false;
    else if (z1s[k] - z0s[k] < VR_SDF_HALO) ok = false;  // a slab thinner than the halo would forward stale planes
  }
  if (!ok) return vrk_sdf_build(ctx, vol, nx, ny, nz, tf, field, levels_out, max_it_out, surf);
  const int z0 = z0s[rank], z1 = z1s[rank];
  const int lo = rank > 0 ? VR_SDF_HALO : 0, hi = rank + 1 < n ? VR_SDF_HALO : 0;
  const int nz_ext = (z1 - z0) + lo + hi, n_own = z1 - z0;
  const size_t plane_vox = (size_t)nx * ny;
  vr_sdf_slab* s = nullptr;
  VR_TRY(vrk_sdf_slab_create(ctx, vol + plane_vox * (z0 - lo), nx, ny, nz_ext, tf, max_it, &s));
  const size_t pw = vrk_sdf_slab_plane_words(s);
  int st = VR_OK;
  while (st == VR_OK && !vrk_sdf_slab_finished(s)) {
    st = vrk_sdf_slab_advance(s, VR_SDF_K, nullptr);
    if (st != VR_OK || vrk_sdf_slab_finished(s)) break;
    uint32_t* bits = vrk_sdf_slab_bits(s);
    ncclResult_t e = ncclGroupStart();
    if (e == ncclSuccess && lo) {  // my lowest interior planes go down; the neighbour's top planes arrive in my lower halo
      e = ncclSend(bits + pw * lo, pw * lo, ncclUint32, rank - 1, comm_of(ctx), ctx->stream);
      if (e == ncclSuccess) e = ncclRecv(bits, pw * lo, ncclUint32, rank - 1, comm_of(ctx), ctx->stream);
    }
    if (e == ncclSuccess && hi) {
      e = ncclSend(bits + pw * (lo + n_own - hi), pw * hi, ncclUint32, rank + 1, comm_of(ctx), ctx->stream);
      if (e == ncclSuccess) e = ncclRecv(bits + pw * (lo + n_own), pw * hi, ncclUint32, rank + 1, comm_of(ctx), ctx->stream);
    }
    const ncclResult_t e2 = ncclGroupEnd();
    if (e == ncclSuccess) e = e2;
    if (e != ncclSuccess) { vr_set_error("vr_sdf_build_sharded: halo swap: %s", ncclGetErrorString(e)); st = VR_ERR_CUDA; break; }
    vrk_sdf_slab_mark_imported(s);
  }
  // slab field -> own brick layers of the global field
  int8_t* slab_field = nullptr;
  if (st == VR_OK) {
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&slab_field), vrk_sdf_field_bytes(nx, ny, nz_ext), ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_sdf_build_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  if (st == VR_OK) st = vrk_sdf_slab_assemble(s, slab_field, 0);
  const size_t layer = (size_t)(nx / BR + 1) * (ny / BR + 1) * BRV;  // bytes of one brick layer
  size_t off[64], len[64];
  const int layers_total = nz / BR + 1;  // incl. the apron layer (or the partial last layer)
  for (int k = 0; k < n; ++k) {
    const int l0 = z0s[k] / BR, l1 = k + 1 < n ? z1s[k] / BR : layers_total;
    off[k] = layer * l0; len[k] = layer * (l1 - l0);
  }
  if (st == VR_OK) {
    cudaError_t e = cudaMemcpyAsync(field + off[rank], slab_field + layer * (lo / BR), len[rank], cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_sdf_build_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  if (st == VR_OK && gather) st = gather_ranges(ctx, field, off, len);
  if (st == VR_OK && gather && surf) st = vrk_sdf_to_surface(ctx, field, nx, ny, nz, surf);
  if (slab_field) cudaFreeAsync(slab_field, ctx->stream);
  if (st == VR_OK) st = vrk_sdf_slab_status(s);  // synchronises the stream (the destroy below would anyway)
  vrk_sdf_slab_destroy(s);
  *levels_out = 0;          // diagnostics only: the sharded build does not collect the per-level change flags
  *max_it_out = max_it;
  return st;
}

extern "C" int vr_sdf_build_sharded(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out) {
  VR_REQUIRE(ctx && vol && out && (rects || n_rects == 0), "vr_sdf_build_sharded: null argument");
  VR_TRY(volume_finish(vol));
  VR_REQUIRE(n_rects >= 0 && n_rects <= VR_TF_MAX_RECTS, "vr_sdf_build_sharded: too many TF clauses");
  return sdf_build_impl(ctx, vol, vr_make_tf_table(rects, n_rects), out, 1);
}
extern "C" int vr_sdf_build_slab_only(vr_ctx* ctx, const vr_volume* vol, const vr_tf_rect* rects, int n_rects, vr_sdf** out) {
  VR_REQUIRE(ctx && vol && out && (rects || n_rects == 0), "vr_sdf_build_slab_only: null argument");
  VR_TRY(volume_finish(vol));
  VR_REQUIRE(n_rects >= 0 && n_rects <= VR_TF_MAX_RECTS, "vr_sdf_build_slab_only: too many TF clauses");
  return sdf_build_impl(ctx, vol, vr_make_tf_table(rects, n_rects), out, 2);
}

extern "C" int vr_renderer_set_sharded_build(vr_renderer* r, int enable) {
  VR_REQUIRE(r, "vr_renderer_set_sharded_build: null argument");
  r->sharded_build = enable != 0;
  return VR_OK;
}

// ---- z-slab histogram and volume filter --------------------------------------------------------------------------------------
extern "C" int vr_histogram_sharded(const vr_volume* v, int width, int height, const float range[4], uint32_t* bins_out) {
  VR_REQUIRE(v && range && bins_out, "vr_histogram_sharded: null argument");
  VR_TRY(volume_finish(v));
  VR_REQUIRE(width > 0 && height > 0 && (size_t)width * height < ((size_t)1 << 31), "vr_histogram_sharded: bad bin grid");
  VR_REQUIRE(v->sampling == VR_SAMPLING_NEAREST, "vr_histogram_sharded: NEAREST sampling only");
  vr_ctx* ctx = v->ctx;
  VR_CUDA(cudaSetDevice(ctx->device));
  int z0 = 0, z1 = v->nz;
  if (ctx->comm && ctx->comm_size > 1) VR_TRY(vr_comm_slab(ctx, v->nz, ctx->comm_rank, &z0, &z1));
  uint32_t* bins = nullptr;
  const size_t nb = (size_t)width * height;
  VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&bins), nb * 4, ctx->stream));
  int st = VR_OK;
  if (z1 > z0) st = vrk_histogram(ctx, v->current(), v->nx, v->ny, v->nz, width, height, range, bins, z0, z1, v->stats[0], v->stats[1]);
  else if (cudaMemsetAsync(bins, 0, nb * 4, ctx->stream) != cudaSuccess) st = VR_ERR_CUDA;
  if (st == VR_OK && ctx->comm && ctx->comm_size > 1) {
    ncclResult_t e = ncclAllReduce(bins, bins, nb, ncclUint32, ncclSum, comm_of(ctx), ctx->stream);
    if (e != ncclSuccess) { vr_set_error("vr_histogram_sharded: %s", ncclGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  if (st == VR_OK) {
    cudaError_t e = cudaMemcpyAsync(bins_out, bins, nb * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_histogram_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
  }
  cudaFreeAsync(bins, ctx->stream);
  return st;
}

extern "C" int vr_volume_filter_sharded(vr_volume* v) {
  VR_REQUIRE(v, "vr_volume_filter_sharded: null argument");
  VR_TRY(volume_finish(v));
  VR_REQUIRE(v->sampling == VR_SAMPLING_NEAREST, "vr_volume_filter_sharded: NEAREST sampling only");
  vr_ctx* ctx = v->ctx;
  if (!ctx->comm || ctx->comm_size == 1) return vr_volume_filter(v);
  VR_CUDA(cudaSetDevice(ctx->device));
  const int n = ctx->comm_size, rank = ctx->comm_rank;
  VR_REQUIRE(n <= 64, "vr_volume_filter_sharded: more than 64 ranks");
  const size_t plane_vox = (size_t)v->nx * v->ny, plane = plane_vox * sizeof(int16_t);
  size_t off[64], len[64];
  int z0 = 0, z1 = 0;
  for (int k = 0; k < n; ++k) {
    int a, b;
    VR_TRY(vr_comm_slab(ctx, v->nz, k, &a, &b));
    off[k] = plane * a; len[k] = plane * (b - a);
    if (k == rank) { z0 = a; z1 = b; }
  }
  int16_t* dst = nullptr;
  VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dst), plane * v->nz, ctx->stream));
  int st = VR_OK;
  if (z1 > z0) {
    // 5^3 taps: two halo planes on each side (none at the global faces, where taps read the border colour); the filtered halo
    // planes are computed from incomplete neighbourhoods and discarded
    const int lo = std::min(2, z0), hi = std::min(2, v->nz - z1);
    const int nz_ext = (z1 - z0) + lo + hi;
    int16_t* tmp = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&tmp), plane * nz_ext, ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vr_volume_filter_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
    if (st == VR_OK) st = vrk_bilateral(ctx, v->current() + plane_vox * (z0 - lo), tmp, v->nx, v->ny, nz_ext);
    if (st == VR_OK) {
      e = cudaMemcpyAsync(dst + plane_vox * z0, tmp + plane_vox * lo, plane * (z1 - z0), cudaMemcpyDeviceToDevice, ctx->stream);
      if (e != cudaSuccess) { vr_set_error("vr_volume_filter_sharded: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
    }
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
  }
  if (st == VR_OK) st = gather_ranges(ctx, dst, off, len);
  if (st == VR_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { vr_set_error("vr_volume_filter_sharded: sync failed"); st = VR_ERR_CUDA; }
  if (st != VR_OK) { cudaFreeAsync(dst, ctx->stream); return st; }
  if (v->cropped) { cudaFreeAsync(v->cropped, ctx->stream); v->cropped = dst; }
  else { cudaFreeAsync(v->original, ctx->stream); v->original = dst; }
  v->generation++;
  return VR_OK;
}
