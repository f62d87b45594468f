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
//   i.e. a level-synchronous BFS over the (clamped) corner-neighbour graph, capped at max_it.
//
// B200 formulation:
//   * A value written during level i is i+1 > i, so it can neither satisfy nor break another voxel's `min == i` test in
//     the same level: the update is hazard-free IN PLACE — one int8 field, no ping-pong.
//   * The field lives in 8x8x8 bricks (512 contiguous bytes) covering coordinates 0..n inclusive (apron of zeros): a 32-byte sector is an 8x4x1 patch and a 128-byte line an
//     8x8x2 slab, which is also what the ray marcher's gathers want (vr_render.cu).
//   * Only bricks next to the wavefront are visited: a brick that finalised a voxel at level i enqueues itself and its
//     26 neighbours (deduplicated with an atomicExch stamp) for level i+1.  Each visit stages the brick's 10^3 halo
//     region in shared memory.  The host never reads anything back inside the loop; levels whose work list is empty
//     cost one empty launch.
//   * create_base_image evaluates the TF 9x per voxel; here each CTA evaluates it once per cell of its 10^3 halo region.
#include "vr_device.cuh"

#define BR 8
#define BRV 512
#define HALO 10
#define SDF_THREADS 128

struct BrickDims {
  int nx, ny, nz;  // voxels
  int bx, by, bz;  // bricks per axis
};

__device__ __forceinline__ size_t brick_voxel_addr(const BrickDims& g, int x, int y, int z) {
  const size_t b = ((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx + (x >> 3);
  return b * BRV + ((z & 7) << 6) + ((y & 7) << 3) + (x & 7);
}

// Which of the 27 bricks around a brick can see a change at local voxel (lx,ly,lz) through their halo: per axis the
// brick itself, plus the lower neighbour when l == 0 and the upper one when l == 7.  Bit index = (oz+1)*9+(oy+1)*3+(ox+1).
__device__ __forceinline__ unsigned touch_mask(int lx, int ly, int lz) {
  const unsigned mx = 2u | (lx == 0 ? 1u : 0u) | (lx == 7 ? 4u : 0u);  // bits over ox = -1,0,1
  const unsigned my = 2u | (ly == 0 ? 1u : 0u) | (ly == 7 ? 4u : 0u);
  const unsigned mz = 2u | (lz == 0 ? 1u : 0u) | (lz == 7 ? 4u : 0u);
  unsigned row = 0;  // 9 bits: oy x ox
  if (my & 1u) row |= mx;
  if (my & 2u) row |= mx << 3;
  if (my & 4u) row |= mx << 6;
  unsigned m = 0;
  if (mz & 1u) m |= row;
  if (mz & 2u) m |= row << 9;
  if (mz & 4u) m |= row << 18;
  return m;
}

// enqueue the bricks selected by `mask` (see touch_mask) around brick (bx,by,bz) for `level`, once each
__device__ __forceinline__ void enqueue_neighbourhood(const BrickDims& g, int bx, int by, int bz, int level,
                                                      int* __restrict__ stamp, uint32_t* __restrict__ list,
                                                      unsigned* __restrict__ count, int lane27, unsigned mask) {
  if (lane27 >= 27 || !((mask >> lane27) & 1u)) return;
  const int ox = lane27 % 3 - 1, oy = (lane27 / 3) % 3 - 1, oz = lane27 / 9 - 1;
  const int x = bx + ox, y = by + oy, z = bz + oz;
  if ((unsigned)x >= (unsigned)g.bx || (unsigned)y >= (unsigned)g.by || (unsigned)z >= (unsigned)g.bz) return;
  const uint32_t b = ((uint32_t)z * g.by + y) * g.bx + x;
  if (atomicExch(stamp + b, level) != level) list[atomicAdd(count, 1u)] = b;
}

// ---- create_base_image, signed_distance_field.cl:6-54 — one CTA per brick ---------------------------------------------
__global__ void __launch_bounds__(SDF_THREADS) k_sdf_base(VolView vol, TfTable tf, BrickDims g, int max_it,
                                                          int8_t* __restrict__ field, int* __restrict__ stamp,
                                                          uint32_t* __restrict__ list, unsigned* __restrict__ count) {
  __shared__ uint8_t ev[HALO * HALO * HALO];
  const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  const int x0 = bx * BR - 1, y0 = by * BR - 1, z0 = bz * BR - 1;
  // event state of every cell of the halo region; coordinates clamped into the volume exactly like
  // clamp(offset + location, 0, size-1) does for each corner (signed_distance_field.cl:35)
  for (int i = threadIdx.x; i < HALO * HALO * HALO; i += SDF_THREADS) {
    const int lx = i % HALO, ly = (i / HALO) % HALO, lz = i / (HALO * HALO);
    const int x = min(max(x0 + lx, 0), g.nx - 1), y = min(max(y0 + ly, 0), g.ny - 1), z = min(max(z0 + lz, 0), g.nz - 1);
    ev[i] = voxel_event(vol, tf, x, y, z) != 0;
  }
  __syncthreads();
  unsigned band = 0;
  int8_t* out = field + (((size_t)bz * g.by + by) * g.bx + bx) * BRV;
#pragma unroll
  for (int k = 0; k < BRV / SDF_THREADS; ++k) {
    const int v = threadIdx.x + k * SDF_THREADS;
    const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
    const int x = bx * BR + lx, y = by * BR + ly, z = bz * BR + lz;
    int val = 0;  // apron / padding cells: the border colour 0 (vr_device.cuh SdfView)
    if (x < g.nx && y < g.ny && z < g.nz) {
      const int c = (lz + 1) * HALO * HALO + (ly + 1) * HALO + (lx + 1);
      const int e = ev[c];
      bool homog = true;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int o = ((q & 1) ? 1 : -1) + ((q & 2) ? HALO : -HALO) + ((q & 4) ? HALO * HALO : -HALO * HALO);
        homog &= (ev[c + o] == e);
      }
      val = e ? -1 : 1;
      if (homog) val *= max_it;
      else band |= touch_mask(lx, ly, lz);
    }
    out[v] = (int8_t)val;
  }
  __shared__ unsigned s_mask;
  if (threadIdx.x == 0) s_mask = 0;
  __syncthreads();
  band = __reduce_or_sync(0xffffffffu, band);
  if ((threadIdx.x & 31) == 0 && band) atomicOr(&s_mask, band);
  __syncthreads();
  if (s_mask && max_it > 2) enqueue_neighbourhood(g, bx, by, bz, 1, stamp, list, count, threadIdx.x, s_mask);
}

// ---- one BFS level over the active bricks, in place: create_signed_distance_field, signed_distance_field.cl:89-112 ----
__global__ void __launch_bounds__(SDF_THREADS) k_sdf_level(BrickDims g, int iteration, int max_it,
                                                           int8_t* __restrict__ field, int* __restrict__ stamp,
                                                           const uint32_t* __restrict__ list_in,
                                                           const unsigned* __restrict__ count_in,
                                                           uint32_t* __restrict__ list_out,
                                                           unsigned* __restrict__ count_out) {
  __shared__ int8_t tile[HALO * HALO * HALO];
  __shared__ unsigned s_mask;
  const unsigned n = *count_in;
  for (unsigned j = blockIdx.x; j < n; j += gridDim.x) {
    const uint32_t b = list_in[j];
    const int bx = b % g.bx, by = (b / g.bx) % g.by, bz = b / (g.bx * g.by);
    const int x0 = bx * BR - 1, y0 = by * BR - 1, z0 = bz * BR - 1;
    int8_t* mine = field + (size_t)b * BRV;
    // this thread's 4 voxels first: a brick without candidates needs no halo
    int cur[BRV / SDF_THREADS];
    bool cand = false;
#pragma unroll
    for (int k = 0; k < BRV / SDF_THREADS; ++k) {
      const int v = threadIdx.x + k * SDF_THREADS;
      const int x = bx * BR + (v & 7), y = by * BR + ((v >> 3) & 7), z = bz * BR + (v >> 6);
      cur[k] = mine[v];
      if (x >= g.nx || y >= g.ny || z >= g.nz) cur[k] = 0;  // padding: never a candidate
      cand |= abs(cur[k]) > iteration;
    }
    // the barrier also orders the previous visit's readers of `tile` / `s_mask` before the refill / reset below
    if (!__syncthreads_or(cand)) continue;
    if (threadIdx.x == 0) s_mask = 0;
    for (int i = threadIdx.x; i < HALO * HALO * HALO; i += SDF_THREADS) {
      const int lx = i % HALO, ly = (i / HALO) % HALO, lz = i / (HALO * HALO);
      const int x = min(max(x0 + lx, 0), g.nx - 1), y = min(max(y0 + ly, 0), g.ny - 1), z = min(max(z0 + lz, 0), g.nz - 1);
      tile[i] = field[brick_voxel_addr(g, x, y, z)];
    }
    __syncthreads();
    unsigned changed = 0;
#pragma unroll
    for (int k = 0; k < BRV / SDF_THREADS; ++k) {
      if (abs(cur[k]) <= iteration) continue;
      const int v = threadIdx.x + k * SDF_THREADS;
      const int c = ((v >> 6) + 1) * HALO * HALO + (((v >> 3) & 7) + 1) * HALO + ((v & 7) + 1);
      int nd = 127, abs_added = 0, added = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int o = ((q & 1) ? 1 : -1) + ((q & 2) ? HALO : -HALO) + ((q & 4) ? HALO * HALO : -HALO * HALO);
        const int val = tile[c + o];
        const int a = abs(val);
        abs_added += a;
        added += val;
        nd = min(nd, a);
      }
      if (abs(added) != abs_added) nd = 0;  // corners of mixed sign (signed_distance_field.cl:83-86)
      if (nd != 0 && nd == iteration && iteration + 1 < max_it) {
        mine[v] = (int8_t)(cur[k] < 0 ? -(iteration + 1) : (iteration + 1));
        changed |= touch_mask(v & 7, (v >> 3) & 7, v >> 6);
      }
    }
    changed = __reduce_or_sync(0xffffffffu, changed);
    if ((threadIdx.x & 31) == 0 && changed) atomicOr(&s_mask, changed);
    __syncthreads();
    if (s_mask && iteration + 2 < max_it)
      enqueue_neighbourhood(g, bx, by, bz, iteration + 1, stamp, list_out, count_out, threadIdx.x, s_mask);
  }
}

// bricked -> x-fastest linear (vr_sdf_download; tests/sdf/sdf_test.cpp:24-31 order)
__global__ void __launch_bounds__(256) k_sdf_unbrick(BrickDims g, const int8_t* __restrict__ field,
                                                     int8_t* __restrict__ linear) {
  const size_t n = (size_t)g.nx * g.ny * g.nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % g.nx);
    const size_t t = i / g.nx;
    const int y = (int)(t % g.ny), z = (int)(t / g.ny);
    linear[i] = field[brick_voxel_addr(g, x, y, z)];
  }
}

size_t vrk_sdf_field_bytes(int nx, int ny, int nz) {
  return (size_t)(nx / BR + 1) * (ny / BR + 1) * (nz / BR + 1) * BRV;
}

int vrk_sdf_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field,
                  int* levels_out, int* max_it_out) {
  const int max_it = std::min(std::max(nx, std::max(ny, nz)) / 2, 127);  // signed_distance_field.cpp:11
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t nbricks = (size_t)g.bx * g.by * g.bz;
  // scratch: stamp[nbricks] | list A[nbricks] | list B[nbricks] | counts[130]
  uint32_t* scratch = nullptr;
  const size_t words = nbricks * 3 + 130;
  VR_CUDA(cudaMallocAsync(&scratch, words * 4, ctx->stream));
  VR_CUDA(cudaMemsetAsync(scratch, 0, words * 4, ctx->stream));
  int* stamp = reinterpret_cast<int*>(scratch);
  uint32_t* lists[2] = {scratch + nbricks, scratch + 2 * nbricks};
  unsigned* counts = scratch + 3 * nbricks;
  VolView v{vol, nx, ny, nz};
  k_sdf_base<<<dim3(g.bx, g.by, g.bz), SDF_THREADS, 0, ctx->stream>>>(v, tf, g, max_it, field, stamp, lists[1], counts + 1);
  ctx->launches++;
  // level i finalises magnitude i+1, which is only stored when i+1 < max_it
  const unsigned grid = (unsigned)std::min<size_t>(nbricks, (size_t)ctx->sm_count * 12);
  for (int it = 1; it + 1 < max_it; ++it) {
    k_sdf_level<<<grid, SDF_THREADS, 0, ctx->stream>>>(g, it, max_it, field, stamp, lists[it & 1], counts + it,
                                                      lists[(it + 1) & 1], counts + it + 1);
    ctx->launches++;
  }
  VR_CUDA(cudaGetLastError());
  unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
  VR_CUDA(cudaMemcpyAsync(hc, counts, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  int levels = 0;
  for (int it = 1; it + 1 < max_it; ++it)
    if (hc[it] != 0) levels = it;
  *levels_out = levels;
  *max_it_out = max_it;
  return VR_OK;
}

int vrk_sdf_unbrick(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, int8_t* linear) {
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t n = (size_t)nx * ny * nz;
  const unsigned blocks = (unsigned)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 16);
  k_sdf_unbrick<<<blocks, 256, 0, ctx->stream>>>(g, field, linear);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
