// vr_sdf.cu — signed distance field build.
//
// Reference: opencl_kernels/signed_distance_field.cl (create_base_image :6-54, neightbour_distance_calc :56-87,
// create_signed_distance_field :89-112) driven by app/signed_distance_field.cpp:7-35 — up to 129 full-volume
// ping-pong passes, each with two blocking 4-byte PCIe transfers.
//
// What the reference computes (DESIGN.md §4.2 has the argument; tests/test_parity_gpu.py pins it bit-exactly against
// the reference's own golden vector and the oracle's literal ping-pong restatement):
//   * event(v)  = is_event_gen(volume[v], |grad(v)|)
//   * base(v)   = s(v) * 1 on the event boundary band, s(v) * max_it where all 8 clamped corner neighbours
//                 share v's event state;  s(v) = -1 inside an event, +1 outside
//   * level i   : a voxel still at max_it whose smallest corner magnitude equals i becomes i+1 (sign kept)
//   i.e. a level-synchronous BFS over the (clamped) corner-neighbour graph, capped at max_it:
//   |F(v)| = 1 on the band, else min(max_it, 1 + length of the shortest clamped-corner-step walk from v to the band).
//
// B200 formulation — a bit-parallel wavefront:
//   A voxel's level is the first k at which it belongs to R_k, where R_0 = band and R_k = R_{k-1} | dilate(R_{k-1}); dilate takes
//   the union over the 8 clamped corner offsets and is separable per axis.  With one BIT per voxel (32 voxels along x per
//   word) a level is a handful of shifts and ORs per word and the working set (two 16 MiB bit volumes at 512^3) lives in L2.
//     k_sdf_events   : E = event bit of every voxel (is_event_gen once per voxel; the reference evaluates it 9x)
//     k_sdf_band_bits: R_0 = voxels with a clamped corner of the other event state (signed_distance_field.cl:22-48)
//     k_sdf_wave5    : one launch per level; warp tiles of 128 x 8 x 8 voxels whose 3x3x3 tile neighbourhood did not change in
//                      the previous level are skipped (both bit volumes already agree there).  The level at which a voxel's bit
//                      appears is recorded in 7 bit planes (RED.OR of the new bits into plane j for every set bit j of level+1):
//                      the wave never writes a field byte (ncu on the variants that did: that is where their time went)
//     k_sdf_assemble : planes + event bits -> the bricked int8 field (8x8x8 bricks of 512 bytes, apron of zeros at
//                      x == nx / y == ny / z == nz — what the ray marcher gathers from), written exactly once, coalesced
//   The build is an object that advances level by level (vr_sdf_slab): the single-GPU build runs it to the end, the z-slab
//   sharded build (parallel.py) swaps halo planes of the bit volume between its ranks every K levels.
// Alternative schedules, all bit-exact, live in vr_sdf_variants.cu and are linked into the A/B build only (`make ab` ->
// libvr_ab.so, selected there with VR_SDF_MODE; DESIGN.md §4.2 has the table).
#include <cstring>
#include "vr_sdf_common.cuh"

// ---- one level: R_out = R_in | dilate(R_in) -------------------------------------------------------------------------------
// Same level semantics, one WARP per tile (4 words x 8 rows x 8 planes): lane = lx + 4*ly, the warp walks the planes.  The
// y- and x-dilated rows yd(z') are computed once per plane (10 per tile) and reused by the planes z'-1 and z'+1, and all the
// per-word index arithmetic of k_sdf_wave3 (ncu: ~250 instructions per word, issue-bound) is shared by the 8 words of a
// thread's column: ~40 instructions per word.  No block-level synchronisation.
template <int XW, int TZ>  // words per tile row: lane = lx + XW*ly, tile = XW words x (32/XW) rows x TZ planes
__global__ void __launch_bounds__(256) k_sdf_wave5(WaveDims g, int tx, int ty, int tz, int level,
                                                   const uint32_t* __restrict__ Rin, uint32_t* __restrict__ Rout,
                                                   uint32_t* __restrict__ planes, unsigned nwords, const int* __restrict__ stamp_in,
                                                   int* __restrict__ stamp_out, unsigned* __restrict__ changed_tiles,
                                                   int all_active) {
  constexpr int YR = 32 / XW;
  const int ntiles = tx * ty * tz;
  const unsigned lane = threadIdx.x & 31;
  const int lx = lane & (XW - 1), ly = lane / XW;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; tile < ntiles; tile += nwarps) {
    if (level != 1 && !all_active && stamp_in[tile] != level) continue;  // warp-uniform
    const int ttx = tile % tx, tq = tile / tx;
    const int tty = tq % ty, ttz = tq / ty;
    const int xw = ttx * XW + lx, y = tty * YR + ly, z0 = ttz * TZ;
    const bool xin = xw < g.nxw;
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const uint32_t lm = last ? (1u << g.lastbit) : 0u;  // x == nx-1 is its own +1 neighbour
    const int yc = min(y, g.ny - 1);
    const unsigned plane_stride = (unsigned)g.ny * (unsigned)g.nxw;
    const uint32_t* pm = Rin + (unsigned)max(yc - 1, 0) * (unsigned)g.nxw + (unsigned)min(xw, g.nxw - 1);
    const uint32_t* pp = Rin + (unsigned)min(yc + 1, g.ny - 1) * (unsigned)g.nxw + (unsigned)min(xw, g.nxw - 1);
    const bool lload = lx == 0 && xw > 0 && xin, rload = lx == XW - 1 && xw + 1 < g.nxw;
    uint32_t ydz[TZ + 2];
#pragma unroll
    for (int k = 0; k < TZ + 2; ++k) {
      const unsigned zo = (unsigned)min(max(z0 - 1 + k, 0), g.nz - 1) * plane_stride;
      const uint32_t c0 = xin ? pm[zo] : 0u, c1 = xin ? pp[zo] : 0u;
      uint32_t l0 = __shfl_up_sync(0xffffffffu, c0, 1), r0 = __shfl_down_sync(0xffffffffu, c0, 1);
      uint32_t l1 = __shfl_up_sync(0xffffffffu, c1, 1), r1 = __shfl_down_sync(0xffffffffu, c1, 1);
      if (lx == 0) { l0 = 0u; l1 = 0u; }
      if (lx == XW - 1) { r0 = 0u; r1 = 0u; }
      if (lload) { l0 = (pm - 1)[zo]; l1 = (pp - 1)[zo]; }
      if (rload) { r0 = (pm + 1)[zo]; r1 = (pp + 1)[zo]; }
      if (first) { l0 = c0 << 31; l1 = c1 << 31; }  // x == 0 is its own -1 neighbour
      ydz[k] = __funnelshift_l(l0, c0, 1) | __funnelshift_r(c0, r0, 1) | __funnelshift_l(l1, c1, 1) | __funnelshift_r(c1, r1, 1) |
               ((c0 | c1) & lm);
    }
    bool changed = false;
    if (xin && y < g.ny) {
      const uint32_t vm = valid_mask(g, xw);
      const unsigned lv = (unsigned)level + 1u;
#pragma unroll
      for (int j = 0; j < TZ; ++j) {
        const int z = z0 + j;
        if (z >= g.nz) break;
        const unsigned w = (unsigned)z * plane_stride + (unsigned)y * (unsigned)g.nxw + (unsigned)xw;
        const uint32_t old = Rin[w];
        const uint32_t now = (old | ydz[j] | ydz[j + 2]) & vm;
        Rout[w] = now;
        const uint32_t diff = now & ~old;
        if (diff) {
          changed = true;
#pragma unroll
          for (int b = 0; b < 7; ++b)
            if ((lv >> b) & 1u) atomicOr(planes + (size_t)b * nwords + w, diff);  // result unused: RED.OR
        }
      }
    }
    if (__any_sync(0xffffffffu, changed)) {
      if (lane < 27) {
        const int ox = lane % 3 - 1, oy = (lane / 3) % 3 - 1, oz = lane / 9 - 1;
        const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
        if ((unsigned)ax < (unsigned)tx && (unsigned)ay < (unsigned)ty && (unsigned)az < (unsigned)tz)
          stamp_out[(az * ty + ay) * tx + ax] = level + 1;
      }
      if (lane == 31) changed_tiles[level] = 1u;
    }
  }
}

// bricked -> x-fastest linear (vr_sdf_download; tests/sdf/sdf_test.cpp:24-31 order).  A brick row is 8 contiguous bytes, so a
// thread moves one 8-voxel group with one 8-byte load; the store is 8 bytes too when the linear rows are 8-byte aligned
// (nx % 8 == 0), byte stores otherwise.  Consecutive threads walk along x, then y, then z.
template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_sdf_unbrick(BrickDims g, const int8_t* __restrict__ field,
                                                     int8_t* __restrict__ linear) {
  const unsigned gx = (unsigned)(g.nx + 7) >> 3;
  const size_t ngroups = (size_t)gx * g.ny * g.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += (size_t)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % gx) << 3;
    const size_t t = i / gx;
    const int y = (int)(t % g.ny), z = (int)(t / g.ny);
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(field + brick_voxel_addr(g, x0, y, z)));
    int8_t* dst = linear + ((size_t)z * g.ny + y) * g.nx + x0;
    if (ALIGNED) *reinterpret_cast<uint2*>(dst) = v;
    else
      for (int k = 0; k < 8 && x0 + k < g.nx; ++k) dst[k] = (int8_t)(((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xFF);
  }
}

size_t vrk_sdf_field_bytes(int nx, int ny, int nz) {
  return (size_t)(nx / BR + 1) * (ny / BR + 1) * (nz / BR + 1) * BRV;
}

// ---- z-slab build (multi-GPU): the same kernels on a rank's slab + halo planes, driven level by level ------------
// A rank holds the planes [z0 - h_lo, z1 + h_hi) of the volume (h = 0 at the global faces).  Everything the wave computes for
// a plane depends on the planes within one step per level, so results go stale from the slab's artificial ends inwards by
// one plane per level (two more at the start: gradient taps of the event test and the band test).  The driver
// (cl_volume_renderer_b200/parallel.py) therefore runs K levels, lets the neighbours overwrite the halo planes of the
// current bit volume with their exact interior planes, marks the import (all tiles active for one level) and continues.
struct vr_sdf_slab {
  vr_ctx* ctx = nullptr;
  WaveDims w{};
  int max_it = 0;      // of the GLOBAL volume (signed_distance_field.cpp:11)
  int level = 1;       // next level to run
  size_t nwords = 0, ntiles = 0;
  int tile_z = WT_Z;   // planes per warp tile of k_sdf_wave5
  uint32_t* scratch = nullptr;  // E | R0 | R1 | stamps[2][ntiles] | changed[130]
  uint32_t* planes = nullptr;
  bool all_active = false;
  uint32_t* E() const { return scratch; }
  uint32_t* R(int i) const { return scratch + (1 + i) * nwords; }
  int* stamps(int i) const { return reinterpret_cast<int*>(scratch + 3 * nwords + (size_t)i * ntiles); }
  unsigned* changed() const { return scratch + 3 * nwords + 2 * ntiles; }
};

int vrk_sdf_slab_create(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int max_it, vr_sdf_slab** out) {
  vr_sdf_slab* s = new (std::nothrow) vr_sdf_slab();
  if (!s) return VR_ERR_NOMEM;
  s->ctx = ctx; s->max_it = max_it;
  WaveDims& w = s->w;
  w.nx = nx; w.ny = ny; w.nz = nz;
  w.nxw = (nx + 31) / 32;
  w.bx = nx / BR + 1; w.by = ny / BR + 1; w.bz = nz / BR + 1;
  // measured at 512^3 (VR_SDF_TZ): see DESIGN.md 4.2
#ifdef VR_AB
  static const int tile_z_env = getenv("VR_SDF_TZ") ? atoi(getenv("VR_SDF_TZ")) : WT_Z;
  s->tile_z = (tile_z_env == 2 || tile_z_env == 4 || tile_z_env == 16) ? tile_z_env : WT_Z;
#else
  s->tile_z = WT_Z;
#endif
  w.tx = (w.nxw + WT_XW - 1) / WT_XW; w.ty = (ny + WT_Y - 1) / WT_Y; w.tz = (nz + s->tile_z - 1) / s->tile_z;
  w.lastbit = (unsigned)((nx - 1) & 31);
  s->nwords = (size_t)w.nxw * ny * nz;
  s->ntiles = (size_t)w.tx * w.ty * w.tz;
  const size_t nwords = s->nwords;
  cudaError_t e = cudaMallocAsync(&s->scratch, (3 * nwords + 2 * s->ntiles + 130) * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&s->planes, 7 * nwords * 4, ctx->stream);
  if (e != cudaSuccess) { vr_set_error("vr_sdf_slab_create: %s", cudaGetErrorString(e)); delete s; return VR_ERR_CUDA; }
  VR_CUDA(cudaMemsetAsync(s->stamps(0), 0, (2 * s->ntiles + 130) * 4, ctx->stream));
  VR_CUDA(cudaMemsetAsync(s->planes + nwords, 0, 6 * nwords * 4, ctx->stream));
  VolView v{vol, nx, ny, nz};
  if (!tf.needs_gradient && nx % 8 == 0) {
    const unsigned chunks = (unsigned)div_up(nx, 256);
    const unsigned nitems = chunks * (unsigned)ny * (unsigned)nz;
    k_sdf_events_v8<<<(unsigned)std::min<size_t>(div_up(nitems, 8), (size_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(
        v, tf, w.nxw, s->E(), chunks, nitems);
  } else {
    const unsigned eg = (unsigned)std::min<size_t>(div_up(nwords, 8), (size_t)ctx->sm_count * 16);
    if (tf.needs_gradient) k_sdf_events<true><<<eg, 256, 0, ctx->stream>>>(v, tf, w.nxw, s->E(), (unsigned)nwords);
    else k_sdf_events<false><<<eg, 256, 0, ctx->stream>>>(v, tf, w.nxw, s->E(), (unsigned)nwords);
  }
  const unsigned bb = (unsigned)std::min<size_t>(div_up(nwords, 256), (size_t)ctx->sm_count * 16);
  k_sdf_band_bits<<<bb, 256, 0, ctx->stream>>>(w, s->E(), s->R(0), s->R(1), s->planes, (unsigned)nwords);
  ctx->launches += 2;
  VR_CUDA(cudaGetLastError());
  *out = s;
  return VR_OK;
}

int vrk_sdf_slab_advance(vr_sdf_slab* s, int nlevels, int* done) {
  vr_ctx* ctx = s->ctx;
  const WaveDims& w = s->w;
  // measured at 512^3 (grid multiplier): 8..12 3.9 ms, 16 3.53, 32 (a warp per tile, no loop) 3.43
#ifdef VR_AB
  static const int grid_mult = getenv("VR_SDF_GRID") ? std::max(atoi(getenv("VR_SDF_GRID")), 1) : 64;
  static const int cta_warps = getenv("VR_SDF_WARPS") ? std::min(std::max(atoi(getenv("VR_SDF_WARPS")), 1), 8) : 4;
#else
  const int grid_mult = 64, cta_warps = 4;
#endif
  const unsigned wg5 = (unsigned)std::min<size_t>(div_up(s->ntiles, cta_warps), (size_t)ctx->sm_count * grid_mult * 4 / cta_warps);
  int n = 0;
  for (; n < nlevels && s->level + 1 < s->max_it; ++n, ++s->level) {
    const int it = s->level;
#define VR_WAVE5(TZ)                                                                                                          \
  k_sdf_wave5<4, TZ><<<wg5, 32 * cta_warps, 0, ctx->stream>>>(w, w.tx, w.ty, w.tz, it, s->R((it + 1) & 1), s->R(it & 1), s->planes, \
                                                              (unsigned)s->nwords, s->stamps(it & 1), s->stamps((it + 1) & 1),   \
                                                              s->changed(), s->all_active ? 1 : 0)
    switch (s->tile_z) {
#ifdef VR_AB
      case 2: VR_WAVE5(2); break;
      case 4: VR_WAVE5(4); break;
      case 16: VR_WAVE5(16); break;
#endif
      default: VR_WAVE5(8); break;
    }
#undef VR_WAVE5
    s->all_active = false;
    ctx->launches++;
  }
  VR_CUDA(cudaGetLastError());
  if (done) *done = n;
  return VR_OK;
}

uint32_t* vrk_sdf_slab_bits(vr_sdf_slab* s) { return s->R((s->level + 1) & 1); }  // the volume level `s->level` will read
size_t vrk_sdf_slab_plane_words(const vr_sdf_slab* s) { return (size_t)s->w.nxw * s->w.ny; }
void vrk_sdf_slab_mark_imported(vr_sdf_slab* s) { s->all_active = true; }
int vrk_sdf_slab_level(const vr_sdf_slab* s) { return s->level; }
bool vrk_sdf_slab_finished(const vr_sdf_slab* s) { return s->level + 1 >= s->max_it; }

int vrk_sdf_slab_assemble(vr_sdf_slab* s, int8_t* field, cudaSurfaceObject_t surf) {
  const WaveDims& w = s->w;
  const unsigned nxwf = (unsigned)((8 * w.bx + 31) / 32);
  const unsigned items = nxwf * (unsigned)w.by * (8u * (unsigned)w.bz);
  const unsigned bg = (unsigned)std::min<size_t>(div_up(items, 8), (size_t)s->ctx->sm_count * 16);
  k_sdf_assemble<<<bg, 256, 0, s->ctx->stream>>>(w, s->max_it, s->E(), s->planes, (unsigned)s->nwords, field, nxwf, items, surf);
  s->ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

void vrk_sdf_slab_destroy(vr_sdf_slab* s) {
  if (!s) return;
  cudaFreeAsync(s->scratch, s->ctx->stream);
  cudaFreeAsync(s->planes, s->ctx->stream);
  cudaStreamSynchronize(s->ctx->stream);
  delete s;
}

// bricked field -> 3-D array behind a surface object: one 8-byte brick row per thread
__global__ void __launch_bounds__(256) k_sdf_to_surface(BrickDims g, const int8_t* __restrict__ field, cudaSurfaceObject_t surf) {
  const size_t rows = (size_t)g.bx * g.by * g.bz * 64;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i >> 6;
    const int in = (int)(i & 63);
    const int bx = (int)(b % g.bx), by = (int)((b / g.bx) % g.by), bz = (int)(b / ((size_t)g.bx * g.by));
    const int x0 = bx * 8, y = by * 8 + (in & 7), z = bz * 8 + (in >> 3);
    if (x0 >= g.nx || y >= g.ny || z >= g.nz) continue;
    const uint2 v = *reinterpret_cast<const uint2*>(field + i * 8);
    if (x0 + 8 <= g.nx) surf3Dwrite(v, surf, x0, y, z);
    else
      for (int k = 0; x0 + k < g.nx; ++k)
        surf3Dwrite((signed char)(((k < 4 ? v.x : v.y) >> (8 * (k & 3))) & 0xFF), surf, x0 + k, y, z);
  }
}

int vrk_sdf_to_surface(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, cudaSurfaceObject_t surf) {
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t rows = (size_t)g.bx * g.by * g.bz * 64;
  k_sdf_to_surface<<<(unsigned)std::min<size_t>(div_up(rows, 256), (size_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(g, field, surf);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

#ifdef VR_AB
int vrk_sdf_build_variant(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field, int* levels_out,
                          int* max_it_out);  // vr_sdf_variants.cu: the schedules tried on the way, linked into the A/B build only
#endif

int vrk_sdf_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field, int* levels_out,
                  int* max_it_out, cudaSurfaceObject_t surf) {
#ifdef VR_AB
  static const char* mode = getenv("VR_SDF_MODE");
  if (mode && *mode && strcmp(mode, "default")) {
    VR_TRY(vrk_sdf_build_variant(ctx, vol, nx, ny, nz, tf, field, levels_out, max_it_out));
    return surf ? vrk_sdf_to_surface(ctx, field, nx, ny, nz, surf) : VR_OK;
  }
#endif
  const int max_it = std::min(std::max(nx, std::max(ny, nz)) / 2, 127);  // signed_distance_field.cpp:11
  vr_sdf_slab* s = nullptr;
  VR_TRY(vrk_sdf_slab_create(ctx, vol, nx, ny, nz, tf, max_it, &s));
  int st = vrk_sdf_slab_advance(s, max_it, nullptr);
  if (st == VR_OK) st = vrk_sdf_slab_assemble(s, field, surf);
  // diagnostics: the last level that set a bit
  int levels = 0;
  if (st == VR_OK) {
    unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
    cudaError_t e = cudaMemcpyAsync(hc, s->changed(), sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vrk_sdf_build: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
    for (int it = 1; st == VR_OK && it + 1 < max_it; ++it)
      if (hc[it] != 0) levels = it;
  }
  vrk_sdf_slab_destroy(s);
  *levels_out = levels;
  *max_it_out = max_it;
  return st;
}

int vrk_sdf_unbrick(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, int8_t* linear) {
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t ngroups = (size_t)((nx + 7) / 8) * ny * nz;
  const unsigned blocks = (unsigned)std::min<size_t>(div_up(ngroups, 256), (size_t)ctx->sm_count * 16);
  if (nx % 8 == 0) k_sdf_unbrick<true><<<blocks, 256, 0, ctx->stream>>>(g, field, linear);
  else k_sdf_unbrick<false><<<blocks, 256, 0, ctx->stream>>>(g, field, linear);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
