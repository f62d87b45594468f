// vr_sdf_variants.cu — alternative schedules of the SDF build, selected with VR_SDF_MODE for A/B measurements.
//
// Every variant computes the reference's field bit-exactly (tests/test_parity_gpu.py::test_sdf_alternative_builds_bit_exact);
// vr_sdf.cu holds the production schedule and the description of what is computed, DESIGN.md §4.2 the measured comparison.
//
//   warp   level-synchronous BFS over 8^3 bricks with work lists, one warp per brick visit, in place on the int8 field
//          (a value written during level i is i+1 > i, so the update is hazard-free in place; ncu on its first version: 68 % of
//          the instructions were per-cell address arithmetic of the halo tile load, the row-wise load took 512^3 from 11.4 to
//          7.9 ms)
//   level  the same with one CTA per brick visit
//   async  asynchronous block relaxation (any relaxation order reaches the same fixpoint; re-lowers voxels many times)
//   wave1  bit volumes, one thread per word and level, 8-byte read-modify-write of the field bytes
//   wave2  wave1 + software-pipelined loads over the CTA's tiles, byte stores, no block-level synchronisation
//   wave3  wave1 + bit-sliced level planes instead of field writes (the production kernel's predecessor)
//   wave4  4 levels per launch on a shared-memory tile
//   reg    4 levels per launch, a warp's tile of bits in registers (64-bit cells, shuffles)
//   front  frontier lists of (word, pending bits), atomicOr scatter; falls back to wave1 when a list overflows
#include <cstring>
#include "vr_sdf_common.cuh"

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

// ---- warp: level-synchronous BFS, one WARP per brick -------------------------------------------------------------------------
// Same level semantics as k_sdf_level, but a brick visit is latency-bound (tile load, corner reads, stamp atomics: a handful
// of dependent round trips), so what matters is how many visits are in flight per SM.  One warp per brick with warp-level
// synchronisation only puts 48 visits in flight per SM instead of 12.
#define LEVEL_WARPS 4
__global__ void __launch_bounds__(LEVEL_WARPS * 32) k_sdf_level_warp(BrickDims g, int iteration, int max_it,
                                                                     int8_t* __restrict__ field, int* __restrict__ stamp,
                                                                     const uint32_t* __restrict__ list_in,
                                                                     const unsigned* __restrict__ count_in,
                                                                     uint32_t* __restrict__ list_out,
                                                                     unsigned* __restrict__ count_out) {
  // tile of magnitudes: 100 rows (ty, tz in 0..9) of 16 bytes; cell tx (0..9) of a row lives at byte 3 + tx, so the 8 core
  // cells are two aligned 32-bit words (bytes 4..11) filled from ONE 8-byte load of the owning brick
  __shared__ __align__(16) uint8_t tiles[LEVEL_WARPS][HALO * HALO * 16];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* tile = tiles[warp];
  const unsigned n = *count_in;
  const unsigned nwarps = gridDim.x * LEVEL_WARPS;
  for (unsigned j = blockIdx.x * LEVEL_WARPS + warp; j < n; j += nwarps) {
    const uint32_t b = list_in[j];
    const int bx = b % g.bx, by = (b / g.bx) % g.by, bz = b / (g.bx * g.by);
    const int x0 = bx * BR - 1, y0 = by * BR - 1, z0 = bz * BR - 1;
    const int lox = bx == 0 ? 1 : 0, loy = by == 0 ? 1 : 0, loz = bz == 0 ? 1 : 0;
    const int hix = min(9, g.nx - 1 - x0), hiy = min(9, g.ny - 1 - y0), hiz = min(9, g.nz - 1 - z0);
    const int rx = min(8, g.nx - 1 - x0), ry = min(8, g.ny - 1 - y0), rz = min(8, g.nz - 1 - z0);
    if (rx < 1 || ry < 1 || rz < 1) continue;  // apron-only brick
    int8_t* mine = field + (size_t)b * BRV;
    // candidates first: a brick without voxels above the current level needs no halo
    const int4 own = reinterpret_cast<const int4*>(mine)[lane];  // 16 voxels: lx 0..7 of rows (2*lane, 2*lane+1)
    bool cand = false;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int w = k < 4 ? own.x : (k < 8 ? own.y : (k < 12 ? own.z : own.w));
      const int v = (int)(int8_t)(w >> (8 * (k & 3)));
      const int idx = (int)lane * 16 + k;
      const bool real = (idx & 7) + 1 <= rx && ((idx >> 3) & 7) + 1 <= ry && (idx >> 6) + 1 <= rz;
      cand |= real && abs(v) > iteration;
    }
    if (!__any_sync(0xffffffffu, cand)) continue;
    __syncwarp();
    for (int row = lane; row < HALO * HALO; row += 32) {
      const int ty = row % HALO, tz = row / HALO;
      if (ty < loy || ty > hiy || tz < loz || tz > hiz) continue;  // never read (corner reads are clamped)
      const int y = y0 + ty, z = z0 + tz;
      const size_t rowb = ((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx;  // brick row of this (y,z)
      const unsigned in = ((z & 7) << 6) | ((y & 7) << 3);
      const uint2 core = *reinterpret_cast<const uint2*>(field + (rowb + bx) * BRV + in);
      uint32_t* t32 = reinterpret_cast<uint32_t*>(tile + row * 16);
      t32[1] = __vabs4(core.x);
      t32[2] = __vabs4(core.y);
      if (lox == 0) tile[row * 16 + 3] = (uint8_t)abs((int)field[(rowb + bx - 1) * BRV + in + 7]);
      if (hix == 9) tile[row * 16 + 12] = (uint8_t)abs((int)field[(rowb + bx + 1) * BRV + in]);
    }
    __syncwarp();
    unsigned touched = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int w = k < 4 ? own.x : (k < 8 ? own.y : (k < 12 ? own.z : own.w));
      const int v = (int)(int8_t)(w >> (8 * (k & 3)));
      const int idx = (int)lane * 16 + k;
      const int tx = (idx & 7) + 1, ty = ((idx >> 3) & 7) + 1, tz = (idx >> 6) + 1;
      if (tx > rx || ty > ry || tz > rz || abs(v) <= iteration) continue;
      const int xm = 3 + max(tx - 1, lox), xp = 3 + min(tx + 1, hix);
      const int ym = max(ty - 1, loy) * 16, yp = min(ty + 1, hiy) * 16;
      const int zm = max(tz - 1, loz) * HALO * 16, zp = min(tz + 1, hiz) * HALO * 16;
      // candidates have 8 corners of one sign, so neightbour_distance_calc reduces to the minimum magnitude
      const int nd = min(min(min((int)tile[zm + ym + xm], (int)tile[zm + ym + xp]), min((int)tile[zm + yp + xm], (int)tile[zm + yp + xp])),
                         min(min((int)tile[zp + ym + xm], (int)tile[zp + ym + xp]), min((int)tile[zp + yp + xm], (int)tile[zp + yp + xp])));
      if (nd == iteration && iteration + 1 < max_it) {
        mine[idx] = (int8_t)(v < 0 ? -(iteration + 1) : (iteration + 1));
        touched |= touch_mask(tx - 1, ty - 1, tz - 1);
      }
    }
    touched = __reduce_or_sync(0xffffffffu, touched);
    if (touched && iteration + 2 < max_it)
      enqueue_neighbourhood(g, bx, by, bz, iteration + 1, stamp, list_out, count_out, (int)lane, touched);
  }
}

// ---- asynchronous block relaxation (VR_SDF_MODE=async) -----------------------------------------------------------------------
// The level-synchronous iteration computes, for every voxel outside the band, 1 + the length of the shortest corner-step
// path to the band, capped at max_it.  Shortest-path distances are the unique fixpoint of the relaxation
//     |F(v)|  <-  min(|F(v)|, 1 + min over the 8 clamped corners c of |F(c)|)
// started from the base image (band = 1, everything else = max_it), and ANY order of relaxations reaches it (values only
// ever decrease towards it).  So instead of <= 125 global levels, each visit of a brick relaxes the brick to LOCAL
// convergence against its current halo, and a brick is revisited only when a neighbour changed a voxel its halo can see.
//   * one WARP per brick, the 10^3 halo tile of magnitudes in shared memory (1000 bytes per warp);
//   * every corner step changes z by +-1, so a forward sweep over the planes z = 0..7 (relaxing against plane z-1) followed by
//     a backward sweep (against z+1) propagates along all z-monotone path pieces; pairs of sweeps repeat until one changes
//     nothing, which is the brick's fixpoint for this halo;
//   * corner coordinates are clamped per axis to the volume (signed_distance_field.cl:72) at read time;
//   * band voxels (magnitude 1) can never be lowered (1 + min >= 2), apron cells are never relaxed nor read;
//   * unordered relaxation would lower most voxels many times (first from far-away sources, then from nearer ones), so the
//     rounds are ORDERED like Dial's buckets: round r only accepts values <= limit(r), a window that grows by SDF_WINDOW
//     every SDF_ROUNDS_PER_WINDOW rounds; a candidate above the limit is deferred (the brick re-enqueues itself).  The
//     order only saves work — the fixpoint, hence the result, does not depend on it.
// Rounds run until a round with an unbounded limit enqueues nothing.  tests/test_parity_gpu.py pins the result bit-exactly against the reference's
// golden vector and the oracle's literal level iteration.
#define RELAX_WARPS 4
#define SDF_WINDOW 8
#define SDF_ROUNDS_PER_WINDOW 2
__global__ void __launch_bounds__(RELAX_WARPS * 32) k_sdf_relax(BrickDims g, int round, int limit,
                                                                int8_t* __restrict__ field,
                                                                int* __restrict__ stamp,
                                                                const uint32_t* __restrict__ list_in,
                                                                const unsigned* __restrict__ count_in,
                                                                uint32_t* __restrict__ list_out,
                                                                unsigned* __restrict__ count_out) {
  __shared__ uint8_t tiles[RELAX_WARPS][HALO * HALO * HALO + 24];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* tile = tiles[warp];
  const unsigned n = *count_in;
  const unsigned nwarps = gridDim.x * RELAX_WARPS;
  for (unsigned j = blockIdx.x * RELAX_WARPS + warp; j < n; j += nwarps) {
    const uint32_t b = list_in[j];
    const int bx = b % g.bx, by = (b / g.bx) % g.by, bz = b / (g.bx * g.by);
    const int x0 = bx * BR - 1, y0 = by * BR - 1, z0 = bz * BR - 1;
    // clamp range of corner reads in tile coordinates (tile index = local + 1): [lo, hi] per axis
    const int lox = bx == 0 ? 1 : 0, loy = by == 0 ? 1 : 0, loz = bz == 0 ? 1 : 0;
    const int hix = min(9, g.nx - 1 - x0), hiy = min(9, g.ny - 1 - y0), hiz = min(9, g.nz - 1 - z0);
    // real voxels of this brick: tile indices 1..8 intersected with the volume
    const int rx = min(8, g.nx - 1 - x0), ry = min(8, g.ny - 1 - y0), rz = min(8, g.nz - 1 - z0);
    if (rx < 1 || ry < 1 || rz < 1) continue;  // apron-only brick
    __syncwarp();
    for (int i = lane; i < HALO * HALO * HALO; i += 32) {
      const int lx = i % HALO, ly = (i / HALO) % HALO, lz = i / (HALO * HALO);
      int m = 127;
      if (lx >= lox && lx <= hix && ly >= loy && ly <= hiy && lz >= loz && lz <= hiz)
        m = abs((int)field[brick_voxel_addr(g, x0 + lx, y0 + ly, z0 + lz)]);
      tile[i] = (uint8_t)m;
    }
    __syncwarp();
    // this lane's two columns (lx, ly) of the 8x8 plane
    int cxm[2], cxp[2], cym[2], cyp[2], ctr[2];
    bool real[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = (int)lane + 32 * k;
      const int tx = (idx & 7) + 1, ty = (idx >> 3) + 1;
      real[k] = tx <= rx && ty <= ry;
      cxm[k] = max(tx - 1, lox); cxp[k] = min(tx + 1, hix);
      cym[k] = max(ty - 1, loy) * HALO; cyp[k] = min(ty + 1, hiy) * HALO;
      ctr[k] = ty * HALO + tx;
    }
    bool any_change = false, converged = false, deferred = false;
    for (int pass = 0; pass < 64; ++pass) {
      bool changed = false;
      // forward: plane tz relaxes against plane clamp(tz-1)
      for (int tz = 1; tz <= rz; ++tz) {
        const int pz = max(tz - 1, loz) * HALO * HALO, cz = tz * HALO * HALO;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (!real[k]) continue;
          const int m = min(min((int)tile[pz + cym[k] + cxm[k]], (int)tile[pz + cym[k] + cxp[k]]),
                            min((int)tile[pz + cyp[k] + cxm[k]], (int)tile[pz + cyp[k] + cxp[k]])) + 1;
          if (m < (int)tile[cz + ctr[k]]) {
            if (m <= limit) { tile[cz + ctr[k]] = (uint8_t)m; changed = true; }
            else deferred = true;
          }
        }
        __syncwarp();
      }
      // backward: plane tz relaxes against plane clamp(tz+1)
      for (int tz = rz; tz >= 1; --tz) {
        const int pz = min(tz + 1, hiz) * HALO * HALO, cz = tz * HALO * HALO;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (!real[k]) continue;
          const int m = min(min((int)tile[pz + cym[k] + cxm[k]], (int)tile[pz + cym[k] + cxp[k]]),
                            min((int)tile[pz + cyp[k] + cxm[k]], (int)tile[pz + cyp[k] + cxp[k]])) + 1;
          if (m < (int)tile[cz + ctr[k]]) {
            if (m <= limit) { tile[cz + ctr[k]] = (uint8_t)m; changed = true; }
            else deferred = true;
          }
        }
        __syncwarp();
      }
      if (!__any_sync(0xffffffffu, changed)) { converged = true; break; }
      any_change = true;
    }
    deferred = __any_sync(0xffffffffu, deferred);
    if (!any_change) {
      // nothing could be lowered within the current window; come back when the window has moved
      if (deferred) enqueue_neighbourhood(g, bx, by, bz, round + 1, stamp, list_out, count_out, (int)lane, 1u << 13);
      continue;
    }
    // write back lowered voxels (sign kept), collect which neighbours can see a change
    int8_t* mine = field + (size_t)b * BRV;
    unsigned touched = 0;
    for (int v = lane; v < BRV; v += 32) {
      const int lx = v & 7, ly = (v >> 3) & 7, lz = v >> 6;
      if (lx + 1 > rx || ly + 1 > ry || lz + 1 > rz) continue;
      const int old = mine[v];
      const int m = tile[(lz + 1) * HALO * HALO + (ly + 1) * HALO + (lx + 1)];
      if (m < abs(old)) {
        mine[v] = (int8_t)(old < 0 ? -m : m);
        touched |= touch_mask(lx, ly, lz);
      }
    }
    touched = __reduce_or_sync(0xffffffffu, touched);
    if (converged && !deferred) touched &= ~(1u << 13);  // bit 13 = this brick: at its fixpoint for the current halo
    if (touched) enqueue_neighbourhood(g, bx, by, bz, round + 1, stamp, list_out, count_out, (int)lane, touched);
  }
}

// One warp per (word column, 8 rows of one z): lane = (8-bit piece of the word) * 8 + row, so that the 8 lanes of a piece
// write the 64 contiguous bytes of one z-slice of a brick.
__global__ void __launch_bounds__(256) k_sdf_band(WaveDims g, int max_it, const uint32_t* __restrict__ E,
                                                  uint32_t* __restrict__ Ra, uint32_t* __restrict__ Rb,
                                                  int8_t* __restrict__ field, unsigned nxwf, unsigned items) {
  const unsigned lane = threadIdx.x & 31;
  const int yr = lane & 7, piece = lane >> 3;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned it = warp; it < items; it += nwarps) {
    const unsigned t = it / nxwf;
    const int xw = (int)(it - t * nxwf);
    const int z = (int)(t / (unsigned)g.by), yg = (int)(t - (unsigned)z * (unsigned)g.by);
    const int y = yg * 8 + yr;
    uint32_t own = 0, band = 0, valid = 0;
    if (xw < g.nxw && y < g.ny && z < g.nz) {
      valid = valid_mask(g, xw);
      const bool first = xw == 0, last = xw == g.nxw - 1;
      own = __ldg(E + ((size_t)z * g.ny + y) * g.nxw + xw);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int yy = min(max(y + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(z + ((q & 2) ? 1 : -1), 0), g.nz - 1);
        const uint32_t* row = E + ((size_t)zz * g.ny + yy) * g.nxw;
        const uint32_t c = __ldg(row + xw);
        const uint32_t l = first ? 0u : __ldg(row + xw - 1), r = last ? 0u : __ldg(row + xw + 1);
        band |= (shl_clamped(c, l, first) ^ own) | (shr_clamped(c, r, last, g.lastbit) ^ own);
      }
      band &= valid;
      if (piece == 0) {
        const size_t w = ((size_t)z * g.ny + y) * g.nxw + xw;
        Ra[w] = band;
        Rb[w] = band;
      }
    }
    const int brick_x = xw * 4 + piece;
    if (brick_x < g.bx) {
      const uint32_t e8 = (own >> (8 * piece)) & 0xFFu, b8 = (band >> (8 * piece)) & 0xFFu, v8 = (valid >> (8 * piece)) & 0xFFu;
      // per byte: valid ? (event ? -1 : 1) * (band ? 1 : max_it) : 0
      uint32_t out[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t ev = bits4_to_bytes(e8 >> (4 * h)), bd = bits4_to_bytes(b8 >> (4 * h)), vd = bits4_to_bytes(v8 >> (4 * h));
        const uint32_t mag = bd + (0x01010101u - bd) * (uint32_t)max_it;  // per byte 1 or max_it (<= 127: no carries)
        const uint32_t val = (mag ^ (ev * 0xFFu)) + ev;                    // per byte -m = ~m + 1 (m >= 1: no carry out)
        out[h] = val & (vd * 0xFFu);
      }
      const size_t brick = ((size_t)(z >> 3) * g.by + yg) * g.bx + brick_x;
      *reinterpret_cast<uint2*>(field + brick * BRV + ((z & 7) << 6) + (yr << 3)) = make_uint2(out[0], out[1]);
    }
  }
}

// One level: R_out = R_in | dilate(R_in) on the active tiles; the voxels whose bit appears get +-(level+1).
// Thread = one word: lane = lx + 4*ly (4 words x 8 rows), warp = lz.  The x-neighbour words of the 4 corner rows come from
// the neighbouring lanes by shuffle; only the lanes at the tile's x edges load them.
__global__ void __launch_bounds__(WAVE_THREADS) k_sdf_wave(WaveDims g, int level, const uint32_t* __restrict__ Rin,
                                                           uint32_t* __restrict__ Rout, const uint32_t* __restrict__ E,
                                                           int8_t* __restrict__ field, const int* __restrict__ stamp_in,
                                                           int* __restrict__ stamp_out, unsigned* __restrict__ changed_tiles) {
  const int ntiles = g.tx * g.ty * g.tz;
  const int lx = threadIdx.x & (WT_XW - 1), ly = (threadIdx.x >> 2) & (WT_Y - 1), lz = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (level != 1 && stamp_in[tile] != level) continue;  // uniform per CTA
    const int ttx = tile % g.tx, tq = tile / g.tx;
    const int tty = tq % g.ty, ttz = tq / g.ty;
    const int xw = ttx * WT_XW + lx, y = tty * WT_Y + ly, z = ttz * WT_Z + lz;
    const bool inside = xw < g.nxw && y < g.ny && z < g.nz;
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const int yc = min(y, g.ny - 1), zc = min(z, g.nz - 1);
    uint32_t dil = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = min(max(yc + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(zc + ((q & 2) ? 1 : -1), 0), g.nz - 1);
      const uint32_t* row = Rin + ((unsigned)zz * (unsigned)g.ny + (unsigned)yy) * (unsigned)g.nxw;
      const uint32_t c = xw < g.nxw ? row[xw] : 0u;
      uint32_t l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
      if (lx == 0) l = (xw > 0 && xw <= g.nxw) ? row[xw - 1] : 0u;
      if (lx == WT_XW - 1) r = (xw + 1 < g.nxw) ? row[xw + 1] : 0u;
      dil |= shl_clamped(c, l, first) | shr_clamped(c, r, last, g.lastbit);
    }
    bool changed = false;
    if (inside) {
      const unsigned w = ((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw;
      const uint32_t old = Rin[w];
      const uint32_t now = (old | dil) & valid_mask(g, xw);
      Rout[w] = now;
      const uint32_t diff = now & ~old;
      if (diff) {
        changed = true;
        const uint32_t e = __ldg(E + w);
        int8_t* rowbase = field + (((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx + (size_t)xw * 4) * BRV + ((z & 7) << 6) + ((y & 7) << 3);
        const uint32_t mag = 0x01010101u * (uint32_t)(level + 1);
#pragma unroll
        for (int piece = 0; piece < 4; ++piece) {
          const uint32_t d8 = (diff >> (8 * piece)) & 0xFFu;
          if (!d8) continue;
          const uint32_t e8 = (e >> (8 * piece)) & 0xFFu;
          uint2* ptr = reinterpret_cast<uint2*>(rowbase + piece * BRV);
          uint2 cur = *ptr;
          const uint32_t m0 = bits4_to_bytes(d8) * 0xFFu, m1 = bits4_to_bytes(d8 >> 4) * 0xFFu;
          const uint32_t ev0 = bits4_to_bytes(e8), ev1 = bits4_to_bytes(e8 >> 4);
          cur.x = (cur.x & ~m0) | (((mag ^ (ev0 * 0xFFu)) + ev0) & m0);
          cur.y = (cur.y & ~m1) | (((mag ^ (ev1 * 0xFFu)) + ev1) & m1);
          *ptr = cur;
        }
      }
    }
    if (__syncthreads_or(changed)) {
      if (threadIdx.x < 27) {
        const int ox = threadIdx.x % 3 - 1, oy = (threadIdx.x / 3) % 3 - 1, oz = threadIdx.x / 9 - 1;
        const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
        if ((unsigned)ax < (unsigned)g.tx && (unsigned)ay < (unsigned)g.ty && (unsigned)az < (unsigned)g.tz)
          stamp_out[(az * g.ty + ay) * g.tx + ax] = level + 1;
      }
      if (threadIdx.x == 32) atomicAdd(changed_tiles + level, 1u);
    }
  }
}

__global__ void __launch_bounds__(WAVE_THREADS) k_sdf_wave3(WaveDims g, int level, const uint32_t* __restrict__ Rin,
                                                            uint32_t* __restrict__ Rout, uint32_t* __restrict__ planes,
                                                            unsigned nwords, const int* __restrict__ stamp_in,
                                                            int* __restrict__ stamp_out, unsigned* __restrict__ changed_tiles) {
  const int ntiles = g.tx * g.ty * g.tz;
  const int lx = threadIdx.x & (WT_XW - 1), ly = (threadIdx.x >> 2) & (WT_Y - 1), lz = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (level != 1 && stamp_in[tile] != level) continue;  // uniform per CTA
    const int ttx = tile % g.tx, tq = tile / g.tx;
    const int tty = tq % g.ty, ttz = tq / g.ty;
    const int xw = ttx * WT_XW + lx, y = tty * WT_Y + ly, z = ttz * WT_Z + lz;
    const bool inside = xw < g.nxw && y < g.ny && z < g.nz;
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const int yc = min(y, g.ny - 1), zc = min(z, g.nz - 1);
    const unsigned w = ((unsigned)zc * (unsigned)g.ny + (unsigned)yc) * (unsigned)g.nxw + (unsigned)min(xw, g.nxw - 1);
    const uint32_t old = Rin[w];
    uint32_t dil = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = min(max(yc + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(zc + ((q & 2) ? 1 : -1), 0), g.nz - 1);
      const uint32_t* row = Rin + ((unsigned)zz * (unsigned)g.ny + (unsigned)yy) * (unsigned)g.nxw;
      const uint32_t c = xw < g.nxw ? row[xw] : 0u;
      uint32_t l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
      if (lx == 0) l = (xw > 0 && xw <= g.nxw) ? row[xw - 1] : 0u;
      if (lx == WT_XW - 1) r = (xw + 1 < g.nxw) ? row[xw + 1] : 0u;
      dil |= shl_clamped(c, l, first) | shr_clamped(c, r, last, g.lastbit);
    }
    bool changed = false;
    if (inside) {
      const uint32_t now = (old | dil) & valid_mask(g, xw);
      Rout[w] = now;
      const uint32_t diff = now & ~old;
      if (diff) {
        changed = true;
        const unsigned lv = (unsigned)level + 1u;
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if ((lv >> j) & 1u) atomicOr(planes + (size_t)j * nwords + w, diff);  // result unused: RED.OR
      }
    }
    if (__syncthreads_or(changed)) {
      if (threadIdx.x < 27) {
        const int ox = threadIdx.x % 3 - 1, oy = (threadIdx.x / 3) % 3 - 1, oz = threadIdx.x / 9 - 1;
        const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
        if ((unsigned)ax < (unsigned)g.tx && (unsigned)ay < (unsigned)g.ty && (unsigned)az < (unsigned)g.tz)
          stamp_out[(az * g.ty + ay) * g.tx + ax] = level + 1;
      }
      if (threadIdx.x == 32) changed_tiles[level] = 1u;
    }
  }
}

// ---- wave2: dense per-level kernel with software-pipelined loads -------------------------------------------------------------
// ncu on k_sdf_wave: ~20 % issue-active; a tile visit is a chain of dependent L2 round trips (stamp -> words -> event word ->
// field bytes -> barrier).  Here every load of a visit (stamp, the 4 corner-row words, the x-edge words, the word itself, the
// event word) is issued at once and one visit ahead (software pipelining over the CTA's tiles), the field bytes are plain
// byte stores (no read-modify-write), and the warps of a CTA never synchronise: each warp owns one z-plane of the tile and
// raises the neighbour stamps itself.
struct WaveTileData {
  int stamp;
  uint32_t c[4], edge[4], old, e;
};

__device__ __forceinline__ void wave2_load(const WaveDims& g, int tile, int level, int lx, int ly, int lz,
                                           const uint32_t* __restrict__ Rin, const uint32_t* __restrict__ E,
                                           const int* __restrict__ stamp_in, WaveTileData& d) {
  d.stamp = level == 1 ? level : stamp_in[tile];
  const int ttx = tile % g.tx, tq = tile / g.tx;
  const int tty = tq % g.ty, ttz = tq / g.ty;
  const int xw = ttx * WT_XW + lx, y = tty * WT_Y + ly, z = ttz * WT_Z + lz;
  const int yc = min(y, g.ny - 1), zc = min(z, g.nz - 1);
  const bool xin = xw < g.nxw;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int yy = min(max(yc + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(zc + ((q & 2) ? 1 : -1), 0), g.nz - 1);
    const uint32_t* row = Rin + ((unsigned)zz * (unsigned)g.ny + (unsigned)yy) * (unsigned)g.nxw;
    d.c[q] = xin ? row[xw] : 0u;
    d.edge[q] = 0u;
    if (lx == 0 && xw > 0 && xw <= g.nxw) d.edge[q] = row[xw - 1];
    if (lx == WT_XW - 1 && xw + 1 < g.nxw) d.edge[q] = row[xw + 1];
  }
  const unsigned w = ((unsigned)zc * (unsigned)g.ny + (unsigned)yc) * (unsigned)g.nxw + (unsigned)min(xw, g.nxw - 1);
  d.old = Rin[w];
  d.e = __ldg(E + w);
}

__global__ void __launch_bounds__(WAVE_THREADS) k_sdf_wave2(WaveDims g, int level, const uint32_t* __restrict__ Rin,
                                                            uint32_t* __restrict__ Rout, const uint32_t* __restrict__ E,
                                                            int8_t* __restrict__ field, const int* __restrict__ stamp_in,
                                                            int* __restrict__ stamp_out, unsigned* __restrict__ changed_levels) {
  const int ntiles = g.tx * g.ty * g.tz;
  const int lx = threadIdx.x & (WT_XW - 1), ly = (threadIdx.x >> 2) & (WT_Y - 1), lz = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  if (tile >= ntiles) return;
  WaveTileData cur, nxt;
  wave2_load(g, tile, level, lx, ly, lz, Rin, E, stamp_in, cur);
  for (; tile < ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < ntiles) wave2_load(g, next, level, lx, ly, lz, Rin, E, stamp_in, nxt);
    if (cur.stamp == level) {  // uniform per CTA
      const int ttx = tile % g.tx, tq = tile / g.tx;
      const int tty = tq % g.ty, ttz = tq / g.ty;
      const int xw = ttx * WT_XW + lx, y = tty * WT_Y + ly, z = ttz * WT_Z + lz;
      const bool inside = xw < g.nxw && y < g.ny && z < g.nz;
      const bool first = xw == 0, last = xw == g.nxw - 1;
      uint32_t dil = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t c = cur.c[q];
        uint32_t l = __shfl_up_sync(0xffffffffu, c, 1), r = __shfl_down_sync(0xffffffffu, c, 1);
        if (lx == 0) l = cur.edge[q];
        if (lx == WT_XW - 1) r = cur.edge[q];
        dil |= shl_clamped(c, l, first) | shr_clamped(c, r, last, g.lastbit);
      }
      uint32_t diff = 0;
      if (inside) {
        const uint32_t now = (cur.old | dil) & valid_mask(g, xw);
        Rout[((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw] = now;
        diff = now & ~cur.old;
        if (diff) {
          int8_t* rowbase = field + (((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx + (size_t)xw * 4) * BRV + ((z & 7) << 6) + ((y & 7) << 3);
          uint32_t bits = diff;
          do {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            rowbase[(b >> 3) * BRV + (b & 7)] = (int8_t)(((cur.e >> b) & 1u) ? -(level + 1) : (level + 1));
          } while (bits);
        }
      }
      if (__any_sync(0xffffffffu, diff != 0)) {  // this warp's plane changed: wake the tile and its neighbours for the next level
        if (lane < 27) {
          const int ox = lane % 3 - 1, oy = (lane / 3) % 3 - 1, oz = lane / 9 - 1;
          const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
          if ((unsigned)ax < (unsigned)g.tx && (unsigned)ay < (unsigned)g.ty && (unsigned)az < (unsigned)g.tz)
            stamp_out[(az * g.ty + ay) * g.tx + ax] = level + 1;
        }
        if (lane == 31) changed_levels[level] = 1u;
      }
    }
    cur = nxt;
  }
}

// ---- wave4: temporally blocked wavefront, W4_H levels per launch, the tile's bits in shared memory ---------------------------
// k_sdf_wave is latency-bound: a level touches every word once with a handful of dependent L2 round trips.  Here a CTA loads
// a (4+2) x (32+2H) x (32+2H)-word region of R into shared memory, runs H levels on it (the region's rim goes stale by one
// cell per level, the 4 x 32 x 32 interior stays exact), writes the interior back and the field bytes of the bits that
// appeared.  Per level: pass 1  B = dilate_y(dilate_x(A)), pass 2  A |= B[z-1] | B[z+1]  (clamped at the volume faces).
#define W4_VX 4
#define W4_VY 32
#define W4_VZ 32
#define W4_H 4
#define W4_THREADS 512
template <int H>
__global__ void __launch_bounds__(W4_THREADS, 2) k_sdf_wave_tb(WaveDims g, int tx4, int ty4, int tz4, int launch_idx, int level0,
                                                               int nlev, const uint32_t* __restrict__ Rin,
                                                               uint32_t* __restrict__ Rout, const uint32_t* __restrict__ E,
                                                               int8_t* __restrict__ field, const int* __restrict__ stamp_in,
                                                               int* __restrict__ stamp_out, unsigned* __restrict__ diag) {
  constexpr int RX = W4_VX + 2, RY = W4_VY + 2 * H, RZ = W4_VZ + 2 * H;
  constexpr int PLANE = RX * RY;
  extern __shared__ uint32_t sm[];
  uint32_t* A = sm;
  uint32_t* B = sm + PLANE * RZ;
  uint32_t* Es = B + PLANE * RZ;  // event bits of the interior: W4_VX x W4_VY x W4_VZ
  const int tile = blockIdx.x;
  if (launch_idx != 0 && stamp_in[tile] != launch_idx) return;
  const int ttx = tile % tx4, tq = tile / tx4;
  const int tty = tq % ty4, ttz = tq / ty4;
  const int xwb = ttx * W4_VX - 1, yb = tty * W4_VY - H, zb = ttz * W4_VZ - H;
  for (int i = threadIdx.x; i < PLANE * RZ; i += W4_THREADS) {
    const int rx = i % RX, r = i / RX;
    const int ry = r % RY, rz = r / RY;
    const int xw = xwb + rx, y = yb + ry, z = zb + rz;
    uint32_t v = 0;
    if ((unsigned)xw < (unsigned)g.nxw && (unsigned)y < (unsigned)g.ny && (unsigned)z < (unsigned)g.nz)
      v = Rin[((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw];
    A[i] = v;
  }
  for (int i = threadIdx.x; i < W4_VX * W4_VY * W4_VZ; i += W4_THREADS) {
    const int xw = xwb + 1 + (i & (W4_VX - 1)), y = yb + H + ((i / W4_VX) & (W4_VY - 1)), z = zb + H + i / (W4_VX * W4_VY);
    uint32_t v = 0;
    if (xw < g.nxw && y < g.ny && z < g.nz) v = __ldg(E + ((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw);
    Es[i] = v;
  }
  __syncthreads();
  // this thread's column (rx, ry) and half of the planes
  const int col = threadIdx.x % PLANE, zhalf = threadIdx.x / PLANE;
  const int rx = col % RX, ry = col / RX;
  const int xw = xwb + rx, y = yb + ry;
  const bool active = zhalf < 2 && (unsigned)xw < (unsigned)g.nxw && (unsigned)y < (unsigned)g.ny;
  const bool first = xw == 0, last = xw == g.nxw - 1;
  const int offc = ry * RX + rx;
  const int offm = max(min(max(y - 1, 0), g.ny - 1) - yb, 0) * RX + rx;
  const int offp = min(min(max(y + 1, 0), g.ny - 1) - yb, RY - 1) * RX + rx;
  const int rz0 = zhalf * (RZ / 2), rz1 = rz0 + RZ / 2;
  const bool owner = active && rx >= 1 && rx <= W4_VX && ry >= H && ry < H + W4_VY;
  const uint32_t vmask = active ? valid_mask(g, xw) : 0u;
  bool changed = false;
  int maxlev = 0;
  for (int l = 0; l < nlev; ++l) {
    const int lev = level0 + l;
    if (active) {
      for (int rz = rz0; rz < rz1; ++rz) {
        if ((unsigned)(zb + rz) >= (unsigned)g.nz) continue;
        const int p = rz * PLANE;
        const uint32_t c0 = A[p + offm], c1 = A[p + offp];
        const uint32_t l0 = rx > 0 ? A[p + offm - 1] : 0u, r0 = rx < RX - 1 ? A[p + offm + 1] : 0u;
        const uint32_t l1 = rx > 0 ? A[p + offp - 1] : 0u, r1 = rx < RX - 1 ? A[p + offp + 1] : 0u;
        B[p + offc] = shl_clamped(c0, l0, first) | shr_clamped(c0, r0, last, g.lastbit) | shl_clamped(c1, l1, first) |
                      shr_clamped(c1, r1, last, g.lastbit);
      }
    }
    __syncthreads();
    if (active) {
      for (int rz = rz0; rz < rz1; ++rz) {
        const int z = zb + rz;
        if ((unsigned)z >= (unsigned)g.nz) continue;
        const int p = rz * PLANE + offc;
        const int pm = max(min(max(z - 1, 0), g.nz - 1) - zb, 0) * PLANE + offc;
        const int pp = min(min(max(z + 1, 0), g.nz - 1) - zb, RZ - 1) * PLANE + offc;
        const uint32_t old = A[p];
        const uint32_t now = (old | B[pm] | B[pp]) & vmask;
        A[p] = now;
        if (owner && rz >= H && rz < H + W4_VZ) {
          uint32_t diff = now & ~old;
          if (diff) {
            changed = true;
            maxlev = lev;
            const uint32_t e = Es[((rz - H) * W4_VY + (ry - H)) * W4_VX + (rx - 1)];
            int8_t* rowbase = field + (((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx + (size_t)xw * 4) * BRV + ((z & 7) << 6) + ((y & 7) << 3);
            do {
              const int b = __ffs(diff) - 1;
              diff &= diff - 1;
              rowbase[(b >> 3) * BRV + (b & 7)] = (int8_t)(((e >> b) & 1u) ? -(lev + 1) : (lev + 1));
            } while (diff);
          }
        }
      }
    }
    __syncthreads();
  }
  if (owner) {
    for (int rz = max(rz0, H); rz < min(rz1, H + W4_VZ); ++rz) {
      const int z = zb + rz;
      if (z >= g.nz) break;
      Rout[((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw] = A[rz * PLANE + offc];
    }
  }
  if (__syncthreads_or(changed)) {
    if (threadIdx.x < 27) {
      const int ox = threadIdx.x % 3 - 1, oy = (threadIdx.x / 3) % 3 - 1, oz = threadIdx.x / 9 - 1;
      const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
      if ((unsigned)ax < (unsigned)tx4 && (unsigned)ay < (unsigned)ty4 && (unsigned)az < (unsigned)tz4)
        stamp_out[(az * ty4 + ay) * tx4 + ax] = launch_idx + 1;
    }
    for (int o = 16; o > 0; o >>= 1) maxlev = max(maxlev, __shfl_xor_sync(0xffffffffu, maxlev, o));
    if ((threadIdx.x & 31) == 0 && maxlev) atomicMax(diag, (unsigned)maxlev);
  }
}

// ---- front: frontier wavefront, only the words that gain bits are touched ----------------------------------------------------
// R_k = R_{k-1} | dilate(F_{k-1}) with F_{k-1} = R_{k-1} \ R_{k-2}: bits older than the frontier were dilated in earlier levels.
// A level works on a list of UNIQUE words that have pending bits (P): the owner of a word takes  new = P[w] & ~R[w],  sets
// R[w] |= new, writes the field bytes +-(level+1) of the new bits (the whole warp writes one word's 32 bytes at a time), and
// scatters dilate(new) into the pending words of the <= 12 target words (4 clamped corner rows x {left, centre, right})
// with atomicOr — the clamped corner relation is symmetric, so scattering from the source equals gathering at the target.
// The thread whose atomicOr finds a pending word empty appends it to the next list.  Work is proportional to the voxels
// finalised (~1 % of the volume per level at 512^3), not to the volume.  P is double buffered (a level clears the words it
// owns while it fills the next level's).  If a list overflows the build restarts with the dense per-level kernel.
#define FRONT_THREADS 256
struct FrontCtx {
  WaveDims g;
  uint32_t* R;
  uint32_t* Pnext;
  uint2* out;
  unsigned* count_out;
  unsigned cap;
  unsigned* overflow;
};

// scatter dilate(nb) of word (xw,y,z) into Pnext, append the words that had nothing pending; warp-synchronous
__device__ __forceinline__ void front_scatter(const FrontCtx& f, bool have, uint32_t nb, int xw, int y, int z, unsigned lane) {
  const WaveDims& g = f.g;
  uint32_t tws[4] = {0, 0, 0, 0};
  uint32_t tyz[4] = {0, 0, 0, 0};
  unsigned app = 0;  // bit 3q+k: append target k (0 centre, 1 left, 2 right) of row q
  if (have) {
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const uint32_t c = (shl_clamped(nb, 0u, first) | shr_clamped(nb, 0u, last, g.lastbit)) & valid_mask(g, xw);
    const bool lbit = !first && (nb & 1u), rbit = !last && (nb >> 31);
    // three batches of independent memory operations instead of a dependent chain per target (ncu: the chained version
    // spent 64 % of its stall samples on long-scoreboard waits, 11 % issue-active): addresses, R filters, atomicOrs
    uint32_t tc[4], tl[4], tr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = (q & 1) ? min(y + 1, g.ny - 1) : max(y - 1, 0), zz = (q & 2) ? min(z + 1, g.nz - 1) : max(z - 1, 0);
      const uint32_t tw = ((unsigned)zz * (unsigned)g.ny + (unsigned)yy) * (unsigned)g.nxw + (unsigned)xw;
      tws[q] = tw;
      tyz[q] = (unsigned)yy | ((unsigned)zz << 16);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // filter with the (possibly stale) R: bits already reached need no pending entry
      tc[q] = c ? (c & ~f.R[tws[q]]) : 0u;
      tl[q] = lbit ? (0x80000000u & ~f.R[tws[q] - 1]) : 0u;
      tr[q] = rbit ? (1u & ~f.R[tws[q] + 1]) : 0u;
    }
    uint32_t oc[4], ol[4], orr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      oc[q] = tc[q] ? atomicOr(f.Pnext + tws[q], tc[q]) : 1u;
      ol[q] = tl[q] ? atomicOr(f.Pnext + tws[q] - 1, tl[q]) : 1u;
      orr[q] = tr[q] ? atomicOr(f.Pnext + tws[q] + 1, tr[q]) : 1u;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // whoever finds the pending word empty lists it
      if (oc[q] == 0u) app |= 1u << (3 * q);
      if (ol[q] == 0u) app |= 2u << (3 * q);
      if (orr[q] == 0u) app |= 4u << (3 * q);
    }
  }
  const unsigned cnt = (unsigned)__popc(app);
  unsigned incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(f.count_out, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  unsigned slot = base + incl - cnt;
  while (app) {
    const int k = __ffs(app) - 1;
    app &= app - 1;
    const int q = k / 3, side = k - 3 * q;
    const uint32_t tw = tws[q] + (side == 1 ? 0xFFFFFFFFu : (side == 2 ? 1u : 0u));
    if (slot < f.cap) f.out[slot] = make_uint2(tw, tyz[q]);
    else *f.overflow = 1u;
    ++slot;
  }
}

__global__ void __launch_bounds__(FRONT_THREADS) k_sdf_front(FrontCtx f, int level, uint32_t* __restrict__ Pcur,
                                                             const uint32_t* __restrict__ E, int8_t* __restrict__ field,
                                                             const uint2* __restrict__ in, const unsigned* __restrict__ count_in) {
  const WaveDims& g = f.g;
  const unsigned n = min(*count_in, f.cap);
  const unsigned lane = threadIdx.x & 31;
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += stride) {  // warp-uniform trip count
    const unsigned i = i0 + lane;
    uint32_t nb = 0, e = 0;
    int xw = 0, y = 0, z = 0;
    unsigned long long rowbase = 0;
    if (i < n) {
      const uint2 ent = in[i];
      const uint32_t w = ent.x;
      y = (int)(ent.y & 0xFFFFu); z = (int)(ent.y >> 16);
      xw = (int)(w - ((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw);
      const uint32_t pend = Pcur[w], old = f.R[w];
      e = __ldg(E + w);  // unconditionally: one round trip together with the two loads above
      Pcur[w] = 0u;
      nb = pend & ~old;
      if (nb) {
        f.R[w] = old | nb;
        rowbase = (unsigned long long)(field + (((size_t)(z >> 3) * g.by + (y >> 3)) * g.bx + (size_t)xw * 4) * BRV + ((z & 7) << 6) + ((y & 7) << 3));
      }
    }
    // field bytes: one word (32 voxels along x = 4 brick rows of 8 bytes) per step, lane b writes voxel b
    unsigned m = __ballot_sync(0xffffffffu, nb != 0);
    const unsigned lane_off = (lane >> 3) * BRV + (lane & 7);
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t bits = __shfl_sync(0xffffffffu, nb, src), ev = __shfl_sync(0xffffffffu, e, src);
      const unsigned long long rb = __shfl_sync(0xffffffffu, rowbase, src);
      if ((bits >> lane) & 1u) reinterpret_cast<int8_t*>(rb)[lane_off] = (int8_t)(((ev >> lane) & 1u) ? -(level + 1) : (level + 1));
    }
    front_scatter(f, nb != 0, nb, xw, y, z, lane);
  }
}

// level 0: every band word scatters its bits (nothing to finalise: the band got +-1 from k_sdf_band)
__global__ void __launch_bounds__(FRONT_THREADS) k_sdf_front_seed(FrontCtx f, unsigned nrows) {
  const WaveDims& g = f.g;
  const unsigned lane = threadIdx.x & 31;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  // a warp walks whole rows, 32 words at a time
  for (unsigned row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += nwarps) {
    const int z = (int)(row / (unsigned)g.ny), y = (int)(row - (unsigned)z * (unsigned)g.ny);
    for (int x0 = 0; x0 < g.nxw; x0 += 32) {
      const int xw = x0 + (int)lane;
      const uint32_t bits = xw < g.nxw ? f.R[row * (unsigned)g.nxw + (unsigned)xw] : 0u;
      front_scatter(f, bits != 0, bits, xw, y, z, lane);
    }
  }
}

// ---- reg: register-resident wavefront, H levels per launch, a warp's tile of bits lives in registers --------------------------
// The dense per-level kernel is latency-bound (every level is a handful of dependent L2 round trips per word), the shared-
// memory tile pays 2.3x halo redundancy in instructions, the frontier lists pay atomics per word.  Here a WARP owns a tile
// of one 32-voxel word (x) x 32 rows (y, one per lane) x RG_Z planes (z, in registers).  Each cell is a 64-bit value
// [low 16 bits of the right word | the word | high 16 bits of the left word], so x-dilation is two 64-bit shifts, y-dilation
// two warp shuffles, z-dilation an OR of the neighbouring planes' registers: a level touches no memory at all.  The tile's
// rim goes stale by one cell per level; after H levels the (32-2H) x (RG_Z-2H) interior is written back.  The bit volumes
// of this mode are stored [z][xw][y] (y fastest) so that the lanes' loads coalesce.
#define RG_Z 24
#define RG_H 4
#define RG_WARPS 4

__device__ __noinline__ void sdf_emit_bits(int8_t* __restrict__ field, int bx, int by, int xw, int y, int z, uint32_t diff,
                                           uint32_t e, int mag) {
  int8_t* rowbase = field + (((size_t)(z >> 3) * by + (y >> 3)) * bx + (size_t)xw * 4) * BRV + ((z & 7) << 6) + ((y & 7) << 3);
  do {
    const int b = __ffs(diff) - 1;
    diff &= diff - 1;
    rowbase[(b >> 3) * BRV + (b & 7)] = (int8_t)(((e >> b) & 1u) ? -mag : mag);
  } while (diff);
}

template <int H, bool BORDER>
__device__ __forceinline__ void wave_reg_body(const WaveDims& g, int xw, int y0, int z0, int level0, int nlev,
                                              const uint32_t* __restrict__ Rin, uint32_t* __restrict__ Rout,
                                              const uint32_t* __restrict__ Et, int8_t* __restrict__ field,
                                              unsigned* __restrict__ diag) {
  const unsigned lane = threadIdx.x & 31;
  const int y = y0 + (int)lane;
  const bool yin = (unsigned)y < (unsigned)g.ny;
  const bool first = xw == 0, last = xw == g.nxw - 1;
  uint32_t lo[RG_Z], hi[RG_Z];
#pragma unroll
  for (int k = 0; k < RG_Z; ++k) {
    const int z = z0 + k;
    lo[k] = 0; hi[k] = 0;
    if (yin && (unsigned)z < (unsigned)g.nz) {
      const uint32_t* p = Rin + ((unsigned)z * (unsigned)g.nxw + (unsigned)xw) * (unsigned)g.ny + (unsigned)y;
      const uint32_t C = p[0];
      const uint32_t L = first ? 0u : p[-g.ny];
      const uint32_t Rr = last ? 0u : p[g.ny];
      lo[k] = (C << 16) | (L >> 16);
      hi[k] = (Rr << 16) | (C >> 16);
    }
  }
  // x clamps and the valid-bit mask in the 64-bit cell: word bit b sits at cell bit 16 + b
  uint32_t fm_lo = 0, lm_lo = 0, lm_hi = 0, vm_lo = 0xFFFFFFFFu, vm_hi = 0xFFFFFFFFu;
  if (BORDER) {
    if (first) { fm_lo = 0x10000u; vm_lo = 0xFFFF0000u; }
    if (last) {
      const unsigned pb = 16u + g.lastbit;  // cell bit of x == nx-1
      if (pb < 32u) { lm_lo = 1u << pb; vm_lo &= (2u << pb) - 1u; vm_hi = 0u; }
      else { lm_hi = 1u << (pb - 32u); vm_hi = (pb - 32u == 31u) ? 0xFFFFFFFFu : ((2u << (pb - 32u)) - 1u); }
    }
    if (!yin) { vm_lo = 0u; vm_hi = 0u; }
  }
  const bool ytop = BORDER && y == 0, ybot = BORDER && y == g.ny - 1;
  const bool lane_valid = lane >= (unsigned)H && lane < 32u - (unsigned)H && yin;
  int maxlev = 0;
  // event bits of the interior cells (the sign of the field bytes), loaded up front: a load inside the level loop would stall
  // the warp for a memory round trip at every plane the wavefront crosses
  uint32_t ev[RG_Z - 2 * H];
#pragma unroll
  for (int k = H; k < RG_Z - H; ++k) {
    const int z = z0 + k;
    ev[k - H] = 0;
    if (lane_valid && (unsigned)z < (unsigned)g.nz)
      ev[k - H] = __ldg(Et + ((unsigned)z * (unsigned)g.nxw + (unsigned)xw) * (unsigned)g.ny + (unsigned)y);
  }

  auto ydil = [&](uint32_t vlo, uint32_t vhi, uint32_t& olo, uint32_t& ohi) {
    // x: (V << 1) | (V >> 1) on the 64-bit cell, plus the clamped self-neighbours of x == 0 / x == nx-1
    uint32_t xlo = (vlo << 1) | __funnelshift_r(vlo, vhi, 1);
    uint32_t xhi = __funnelshift_l(vlo, vhi, 1) | (vhi >> 1);
    if (BORDER) { xlo |= vlo & (fm_lo | lm_lo); xhi |= vhi & lm_hi; }
    // y: rows y-1 and y+1 live in the neighbouring lanes (clamped at the volume faces)
    uint32_t ulo = __shfl_up_sync(0xffffffffu, xlo, 1), uhi = __shfl_up_sync(0xffffffffu, xhi, 1);
    uint32_t dlo = __shfl_down_sync(0xffffffffu, xlo, 1), dhi = __shfl_down_sync(0xffffffffu, xhi, 1);
    if (BORDER) {
      if (ytop) { ulo = xlo; uhi = xhi; }
      if (ybot) { dlo = xlo; dhi = xhi; }
    }
    olo = ulo | dlo;
    ohi = uhi | dhi;
  };

  for (int l = 0; l < nlev; ++l) {
    const int lev = level0 + l;
    uint32_t mlo, mhi, clo, chi, plo, phi;  // y-dilated planes k-1, k, k+1 (all from the values before this level)
    ydil(lo[0], hi[0], clo, chi);
    mlo = clo; mhi = chi;
#pragma unroll
    for (int k = 0; k < RG_Z; ++k) {
      const int z = z0 + k;
      if (k + 1 < RG_Z) ydil(lo[k + 1], hi[k + 1], plo, phi);
      else { plo = clo; phi = chi; }
      uint32_t nlo, nhi;
      if (BORDER) {
        const bool zlo_face = z == 0, zhi_face = z == g.nz - 1;
        nlo = (zlo_face ? clo : mlo) | (zhi_face ? clo : plo);
        nhi = (zlo_face ? chi : mhi) | (zhi_face ? chi : phi);
        if ((unsigned)z >= (unsigned)g.nz) { nlo = 0u; nhi = 0u; }
        nlo &= vm_lo; nhi &= vm_hi;
      } else {
        nlo = mlo | plo;
        nhi = mhi | phi;
      }
      const uint32_t olo = lo[k], ohi = hi[k];
      lo[k] = olo | nlo;
      hi[k] = ohi | nhi;
      if (k >= H && k < RG_Z - H) {  // interior plane: record the voxels whose bit appeared
        const uint32_t cold = __funnelshift_r(olo, ohi, 16), cnew = __funnelshift_r(lo[k], hi[k], 16);
        const uint32_t diff = cnew & ~cold;
        if (lane_valid && diff && (!BORDER || (unsigned)z < (unsigned)g.nz)) {
          maxlev = lev;
          sdf_emit_bits(field, g.bx, g.by, xw, y, z, diff, ev[k - H], lev + 1);
        }
      }
      mlo = clo; mhi = chi;
      clo = plo; chi = phi;
    }
  }
  if (lane_valid) {
#pragma unroll
    for (int k = H; k < RG_Z - H; ++k) {
      const int z = z0 + k;
      if ((unsigned)z < (unsigned)g.nz)
        Rout[((unsigned)z * (unsigned)g.nxw + (unsigned)xw) * (unsigned)g.ny + (unsigned)y] = __funnelshift_r(lo[k], hi[k], 16);
    }
  }
  for (int o = 16; o > 0; o >>= 1) maxlev = max(maxlev, __shfl_xor_sync(0xffffffffu, maxlev, o));
  if (lane == 0 && maxlev) atomicMax(diag, (unsigned)maxlev);
}

template <int H>
__global__ void __launch_bounds__(RG_WARPS * 32) k_sdf_wave_reg(WaveDims g, int nty, int ntz, int level0, int nlev,
                                                                const uint32_t* __restrict__ Rin, uint32_t* __restrict__ Rout,
                                                                const uint32_t* __restrict__ Et, int8_t* __restrict__ field,
                                                                unsigned* __restrict__ diag) {
  const unsigned tile = blockIdx.x * RG_WARPS + (threadIdx.x >> 5);
  if (tile >= (unsigned)g.nxw * (unsigned)nty * (unsigned)ntz) return;  // warp-uniform
  const int xw = (int)(tile % (unsigned)g.nxw);
  const unsigned t = tile / (unsigned)g.nxw;
  const int ty = (int)(t % (unsigned)nty), tz = (int)(t / (unsigned)nty);
  const int y0 = ty * (32 - 2 * H) - H, z0 = tz * (RG_Z - 2 * H) - H;
  const bool border = xw == 0 || xw == g.nxw - 1 || y0 < 0 || y0 + 32 > g.ny || z0 < 0 || z0 + RG_Z > g.nz;
  if (border) wave_reg_body<H, true>(g, xw, y0, z0, level0, nlev, Rin, Rout, Et, field, diag);
  else wave_reg_body<H, false>(g, xw, y0, z0, level0, nlev, Rin, Rout, Et, field, diag);
}

// [z][y][xw] -> [z][xw][y]
__global__ void __launch_bounds__(256) k_bits_transpose(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int nxw, int ny,
                                                        unsigned nwords) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
    const unsigned row = i / (unsigned)nxw, xw = i - row * (unsigned)nxw;
    const unsigned z = row / (unsigned)ny, y = row - z * (unsigned)ny;
    out[(z * (unsigned)nxw + xw) * (unsigned)ny + y] = in[i];
  }
}

int vrk_sdf_build_variant(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field,
                          int* levels_out, int* max_it_out) {
  const int max_it = std::min(std::max(nx, std::max(ny, nz)) / 2, 127);  // signed_distance_field.cpp:11
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t nbricks = (size_t)g.bx * g.by * g.bz;
  static const char* mode_env = getenv("VR_SDF_MODE");
  static const bool level_sync = mode_env && !strcmp(mode_env, "level");
  static const bool async_relax = mode_env && !strcmp(mode_env, "async");
  static const bool brick_bfs = mode_env && !strcmp(mode_env, "warp");
  if (!level_sync && !async_relax && !brick_bfs) {
    // the bit-volume variants (wave1, wave2, wave3, wave4, reg, front; an unknown mode name runs wave3)
    WaveDims w{};
    w.nx = nx; w.ny = ny; w.nz = nz;
    w.nxw = (nx + 31) / 32;
    w.bx = g.bx; w.by = g.by; w.bz = g.bz;
    w.tx = (w.nxw + WT_XW - 1) / WT_XW; w.ty = (ny + WT_Y - 1) / WT_Y; w.tz = (nz + WT_Z - 1) / WT_Z;
    w.lastbit = (unsigned)((nx - 1) & 31);
    const size_t nwords = (size_t)w.nxw * ny * nz;
    const size_t ntiles = (size_t)w.tx * w.ty * w.tz;
    // scratch: E | Ra | Rb | stamps[2][ntiles] | changed[130]
    uint32_t* scratch = nullptr;
    const size_t words = 3 * nwords + 2 * ntiles + 130;
    VR_CUDA(cudaMallocAsync(&scratch, words * 4, ctx->stream));
    uint32_t *E = scratch, *R[2] = {scratch + nwords, scratch + 2 * nwords};
    int* stamps[2] = {reinterpret_cast<int*>(scratch + 3 * nwords), reinterpret_cast<int*>(scratch + 3 * nwords + ntiles)};
    unsigned* changed = scratch + 3 * nwords + 2 * ntiles;
    VR_CUDA(cudaMemsetAsync(stamps[0], 0, (2 * ntiles + 130) * 4, ctx->stream));
    VolView v{vol, nx, ny, nz};
    if (!tf.needs_gradient && nx % 8 == 0) {
      const unsigned chunks = (unsigned)div_up(nx, 256);
      const unsigned nitems = chunks * (unsigned)ny * (unsigned)nz;
      const unsigned eg = (unsigned)std::min<size_t>(div_up(nitems, 8), (size_t)ctx->sm_count * 16);
      k_sdf_events_v8<<<eg, 256, 0, ctx->stream>>>(v, tf, w.nxw, E, chunks, nitems);
    } else {
      const unsigned eg = (unsigned)std::min<size_t>(div_up(nwords, 8), (size_t)ctx->sm_count * 16);
      if (tf.needs_gradient) k_sdf_events<true><<<eg, 256, 0, ctx->stream>>>(v, tf, w.nxw, E, (unsigned)nwords);
      else k_sdf_events<false><<<eg, 256, 0, ctx->stream>>>(v, tf, w.nxw, E, (unsigned)nwords);
    }
    const unsigned nxwf = (unsigned)((8 * w.bx + 31) / 32);
    const unsigned band_items = nxwf * (unsigned)w.by * (8u * (unsigned)w.bz);
    const unsigned bg = (unsigned)std::min<size_t>(div_up(band_items, 8), (size_t)ctx->sm_count * 16);
    static const bool planes_mode = !mode_env || !(!strcmp(mode_env, "wave1") || !strcmp(mode_env, "wave2") || !strcmp(mode_env, "wave4") ||
                                                   !strcmp(mode_env, "front") || !strcmp(mode_env, "reg"));
    if (!planes_mode) k_sdf_band<<<bg, 256, 0, ctx->stream>>>(w, max_it, E, R[0], R[1], field, nxwf, band_items);
    ctx->launches += 2;
    static const bool per_level = mode_env && !strcmp(mode_env, "wave1");
    static const bool blocked = mode_env && !strcmp(mode_env, "wave4");
    static const bool frontier = mode_env && !strcmp(mode_env, "front");
    static const bool regtile = mode_env && !strcmp(mode_env, "reg");
    static const bool pipelined = mode_env && !strcmp(mode_env, "wave2");
    if (!per_level && !blocked && !frontier && !regtile && !pipelined) {
      // wave3: CTA tiles, levels recorded in bit planes, field assembled once at the end
      uint32_t* planes = nullptr;
      VR_CUDA(cudaMallocAsync(&planes, 7 * nwords * 4, ctx->stream));
      VR_CUDA(cudaMemsetAsync(planes + nwords, 0, 6 * nwords * 4, ctx->stream));
      const unsigned bb = (unsigned)std::min<size_t>(div_up(nwords, 256), (size_t)ctx->sm_count * 16);
      k_sdf_band_bits<<<bb, 256, 0, ctx->stream>>>(w, E, R[0], R[1], planes, (unsigned)nwords);
      ctx->launches++;
      const unsigned wg = (unsigned)std::min<size_t>(ntiles, (size_t)ctx->sm_count * 8);
      for (int it = 1; it + 1 < max_it; ++it) {
        k_sdf_wave3<<<wg, WAVE_THREADS, 0, ctx->stream>>>(w, it, R[(it + 1) & 1], R[it & 1], planes, (unsigned)nwords, stamps[it & 1],
                                                         stamps[(it + 1) & 1], changed);
        ctx->launches++;
      }
      k_sdf_assemble<<<bg, 256, 0, ctx->stream>>>(w, max_it, E, planes, (unsigned)nwords, field, nxwf, band_items, 0);
      ctx->launches++;
      VR_CUDA(cudaGetLastError());
      unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
      VR_CUDA(cudaMemcpyAsync(hc, changed, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaFreeAsync(planes, ctx->stream));
      VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      int levels = 0;
      for (int it = 1; it + 1 < max_it; ++it)
        if (hc[it] != 0) levels = it;
      *levels_out = levels;
      *max_it_out = max_it;
      return VR_OK;
    }
    if (pipelined) {
      // wave2: dense per-level kernel with pipelined loads
      const unsigned wg = (unsigned)std::min<size_t>(ntiles, (size_t)ctx->sm_count * 8);
      for (int it = 1; it + 1 < max_it; ++it) {
        k_sdf_wave2<<<wg, WAVE_THREADS, 0, ctx->stream>>>(w, it, R[(it + 1) & 1], R[it & 1], E, field, stamps[it & 1],
                                                         stamps[(it + 1) & 1], changed);
        ctx->launches++;
      }
      VR_CUDA(cudaGetLastError());
      unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
      VR_CUDA(cudaMemcpyAsync(hc, changed, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      int levels = 0;
      for (int it = 1; it + 1 < max_it; ++it)
        if (hc[it] != 0) levels = it;
      *levels_out = levels;
      *max_it_out = max_it;
      return VR_OK;
    }
    if (regtile) {
      // reg: register-resident tiles, RG_H levels per launch, bit volumes transposed to [z][xw][y]
      constexpr int H = RG_H;
      uint32_t* extra = nullptr;
      VR_CUDA(cudaMallocAsync(&extra, 2 * nwords * 4, ctx->stream));
      uint32_t* T[2] = {R[1], extra};   // R[1] (a copy of the band) is free in this mode
      uint32_t* Et = extra + nwords;
      const unsigned tg = (unsigned)std::min<size_t>(div_up(nwords, 256), (size_t)ctx->sm_count * 16);
      k_bits_transpose<<<tg, 256, 0, ctx->stream>>>(R[0], T[0], w.nxw, ny, (unsigned)nwords);
      k_bits_transpose<<<tg, 256, 0, ctx->stream>>>(E, Et, w.nxw, ny, (unsigned)nwords);
      ctx->launches += 2;
      const int nty = (ny + (32 - 2 * H) - 1) / (32 - 2 * H), ntz = (nz + (RG_Z - 2 * H) - 1) / (RG_Z - 2 * H);
      const unsigned tiles = (unsigned)w.nxw * (unsigned)nty * (unsigned)ntz;
      int launch = 0;
      for (int it = 1; it + 1 < max_it; it += H, ++launch) {
        const int nlev = std::min(H, max_it - 1 - it);
        k_sdf_wave_reg<H><<<div_up(tiles, RG_WARPS), RG_WARPS * 32, 0, ctx->stream>>>(w, nty, ntz, it, nlev, T[launch & 1],
                                                                                   T[(launch + 1) & 1], Et, field, changed);
        ctx->launches++;
      }
      VR_CUDA(cudaGetLastError());
      unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
      VR_CUDA(cudaMemcpyAsync(hc, changed, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaFreeAsync(extra, ctx->stream));
      VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      *levels_out = (int)hc[0];
      *max_it_out = max_it;
      return VR_OK;
    }
    if (frontier && max_it > 2) {
      // frontier lists: 2 x cap entries of (word, y | z << 16) | counts[130] | overflow; pending bits: R[1] and E-sized P2
      static const char* cap_env = getenv("VR_SDF_FRONT_CAP");  // tests force the overflow fallback with a tiny capacity
      const size_t cap = cap_env ? (size_t)atol(cap_env) : nwords + 1024;
      uint32_t* fs = nullptr;
      VR_CUDA(cudaMallocAsync(&fs, (4 * cap + 132 + nwords) * 4, ctx->stream));
      uint2* lists[2] = {reinterpret_cast<uint2*>(fs), reinterpret_cast<uint2*>(fs + 2 * cap)};
      unsigned* counts = fs + 4 * cap;
      unsigned* overflow = counts + 130;
      uint32_t* P[2] = {R[1], fs + 4 * cap + 132};  // R[1] is free in this mode
      VR_CUDA(cudaMemsetAsync(counts, 0, (132 + nwords) * 4, ctx->stream));
      VR_CUDA(cudaMemsetAsync(P[0], 0, nwords * 4, ctx->stream));
      FrontCtx f{w, R[0], P[1], lists[1], counts + 1, (unsigned)cap, overflow};
      const unsigned nrows = (unsigned)ny * (unsigned)nz;
      const unsigned sg = (unsigned)std::min<size_t>(div_up(nrows, FRONT_THREADS / 32), (size_t)ctx->sm_count * 8);
      k_sdf_front_seed<<<sg, FRONT_THREADS, 0, ctx->stream>>>(f, nrows);
      ctx->launches++;
      const unsigned fg = (unsigned)ctx->sm_count * 8;
      for (int it = 1; it + 1 < max_it; ++it) {
        f.Pnext = P[(it + 1) & 1];
        f.out = lists[(it + 1) & 1];
        f.count_out = counts + it + 1;
        k_sdf_front<<<fg, FRONT_THREADS, 0, ctx->stream>>>(f, it, P[it & 1], E, field, lists[it & 1], counts + it);
        ctx->launches++;
      }
      VR_CUDA(cudaGetLastError());
      unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
      VR_CUDA(cudaMemcpyAsync(hc, counts, sizeof(unsigned) * 132, cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaFreeAsync(fs, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      if (hc[130] == 0) {
        int levels = 0;
        for (int it = 1; it + 1 < max_it; ++it)
          if (hc[it] != 0) levels = it;  // words with pending bits at level it
        VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
        *levels_out = levels;
        *max_it_out = max_it;
        return VR_OK;
      }
      // overflow: rebuild R_0 and the field, then run the dense per-level kernel below
      k_sdf_band<<<bg, 256, 0, ctx->stream>>>(w, max_it, E, R[0], R[1], field, nxwf, band_items);
      ctx->launches++;
    }
    if (!blocked) {
      const unsigned wg = (unsigned)std::min<size_t>(ntiles, (size_t)ctx->sm_count * 8);
      for (int it = 1; it + 1 < max_it; ++it) {  // level it finalises magnitude it+1, stored only when it+1 < max_it
        k_sdf_wave<<<wg, WAVE_THREADS, 0, ctx->stream>>>(w, it, R[(it + 1) & 1], R[it & 1], E, field, stamps[it & 1],
                                                        stamps[(it + 1) & 1], changed);
        ctx->launches++;
      }
    } else {
      constexpr int H = W4_H;
      const int tx4 = (w.nxw + W4_VX - 1) / W4_VX, ty4 = (ny + W4_VY - 1) / W4_VY, tz4 = (nz + W4_VZ - 1) / W4_VZ;
      const size_t smem = ((size_t)2 * (W4_VX + 2) * (W4_VY + 2 * H) * (W4_VZ + 2 * H) + W4_VX * W4_VY * W4_VZ) * 4;
      static bool attr_set = false;
      if (!attr_set) {
        VR_CUDA(cudaFuncSetAttribute(k_sdf_wave_tb<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      int launch = 0;
      for (int it = 1; it + 1 < max_it; it += H, ++launch) {
        const int nlev = std::min(H, max_it - 1 - it);
        k_sdf_wave_tb<H><<<tx4 * ty4 * tz4, W4_THREADS, smem, ctx->stream>>>(w, tx4, ty4, tz4, launch, it, nlev, R[launch & 1],
                                                                          R[(launch + 1) & 1], E, field, stamps[launch & 1],
                                                                          stamps[(launch + 1) & 1], changed);
        ctx->launches++;
      }
    }
    VR_CUDA(cudaGetLastError());
    unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
    VR_CUDA(cudaMemcpyAsync(hc, changed, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    int levels = (int)hc[0];  // temporally blocked build: atomicMax of the last level that set a bit
    for (int it = 1; it + 1 < max_it; ++it)
      if (hc[it] != 0) levels = it;
    *levels_out = levels;
    *max_it_out = max_it;
    return VR_OK;
  }
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
  unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
  int levels = 0;
  if (level_sync) {
    // level i finalises magnitude i+1, which is only stored when i+1 < max_it
    const unsigned grid = (unsigned)std::min<size_t>(nbricks, (size_t)ctx->sm_count * 12);
    for (int it = 1; it + 1 < max_it; ++it) {
      k_sdf_level<<<grid, SDF_THREADS, 0, ctx->stream>>>(g, it, max_it, field, stamp, lists[it & 1], counts + it,
                                                        lists[(it + 1) & 1], counts + it + 1);
      ctx->launches++;
    }
    VR_CUDA(cudaGetLastError());
    VR_CUDA(cudaMemcpyAsync(hc, counts, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int it = 1; it + 1 < max_it; ++it)
      if (hc[it] != 0) levels = it;
  } else if (!async_relax) {
    // warp: level-synchronous, one warp per brick
    const unsigned grid = (unsigned)std::min<size_t>(div_up(nbricks, LEVEL_WARPS), (size_t)ctx->sm_count * 12);
    for (int it = 1; it + 1 < max_it; ++it) {
      k_sdf_level_warp<<<grid, LEVEL_WARPS * 32, 0, ctx->stream>>>(g, it, max_it, field, stamp, lists[it & 1], counts + it,
                                                                  lists[(it + 1) & 1], counts + it + 1);
      ctx->launches++;
    }
    VR_CUDA(cudaGetLastError());
    VR_CUDA(cudaMemcpyAsync(hc, counts, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
    VR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int it = 1; it + 1 < max_it; ++it)
      if (hc[it] != 0) levels = it;
  } else if (max_it > 2) {
    // asynchronous block relaxation: rounds until a round enqueues nothing.  counts[] is reused cyclically: slot r % 128
    // is zeroed two rounds before it is written again.
    const unsigned grid = (unsigned)std::min<size_t>(div_up(nbricks, RELAX_WARPS), (size_t)ctx->sm_count * 8);
    static const int win = getenv("VR_SDF_WINDOW") ? atoi(getenv("VR_SDF_WINDOW")) : SDF_WINDOW;
    static const int rpw = getenv("VR_SDF_RPW") ? atoi(getenv("VR_SDF_RPW")) : SDF_ROUNDS_PER_WINDOW;
    const int ordered_rounds = ((max_it + win - 1) / win) * rpw;  // after these the limit is unbounded
    for (int round = 1;;) {
      const int batch = round == 1 ? ordered_rounds + 2 : 8;
      for (int k = 0; k < batch; ++k, ++round) {
        unsigned* cin = counts + (round & 127);
        unsigned* cout = counts + ((round + 1) & 127);
        const int limit = round <= ordered_rounds ? win * ((round + rpw - 1) / rpw) : 127;
        VR_CUDA(cudaMemsetAsync(counts + ((round + 2) & 127), 0, sizeof(unsigned), ctx->stream));
        k_sdf_relax<<<grid, RELAX_WARPS * 32, 0, ctx->stream>>>(g, round, limit, field, stamp, lists[round & 1], cin,
                                                               lists[(round + 1) & 1], cout);
        ctx->launches++;
      }
      VR_CUDA(cudaGetLastError());
      VR_CUDA(cudaMemcpyAsync(hc, counts + (round & 127), sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
      VR_CUDA(cudaStreamSynchronize(ctx->stream));
      levels = round - 1;
      if (hc[0] == 0) break;
    }
  }
  VR_CUDA(cudaFreeAsync(scratch, ctx->stream));
  *levels_out = levels;
  *max_it_out = max_it;
  return VR_OK;
}

