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
// B200 formulation (default build = k_sdf_base + k_sdf_level_warp):
//   * A value written during level i is i+1 > i, so it can neither satisfy nor break another voxel's `min == i` test in
//     the same level: the update is hazard-free IN PLACE — one int8 field, no ping-pong.
//   * The field lives in 8x8x8 bricks (512 contiguous bytes) covering coordinates 0..n inclusive (apron of zeros): a
//     32-byte sector is an 8x4x1 patch and a 128-byte line an 8x8x2 slab, which is also what the ray marcher's gathers
//     want (vr_render.cu).
//   * Only bricks next to the wavefront are visited: a brick that finalised a voxel at level i enqueues itself and those
//     of its 26 neighbours whose halo can see a changed voxel (deduplicated with an atomicExch stamp) for level i+1.
//     The host never reads anything back inside the loop; levels whose work list is empty cost one empty launch.
//   * One WARP per brick visit: 16 voxels per lane, the brick's own 512 bytes as one 16-byte load per lane, the 10^3 halo
//     tile of magnitudes in shared memory filled row-wise (one 8-byte load + two halo bytes per row), warp-level
//     synchronisation only.  ncu showed the first version spending 68 % of its instructions on per-cell address
//     arithmetic of the tile load; the row-wise load took 512^3 from 11.4 to 7.9 ms.
//   * create_base_image evaluates the TF 9x per voxel; here each CTA evaluates it once per cell of its 10^3 halo region.
// Kept for A/B (VR_SDF_MODE=level | async): the CTA-per-brick level kernel and an asynchronous block-relaxation solver
// (any relaxation order reaches the same fixpoint; it is bit-exact too, but re-lowers voxels many times and loses).
#include <cstring>
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

// ---- level-synchronous BFS, one WARP per brick (default build) ------------------------------------------------------------
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
  static const char* mode_env = getenv("VR_SDF_MODE");
  static const bool level_sync = mode_env && !strcmp(mode_env, "level");
  static const bool async_relax = mode_env && !strcmp(mode_env, "async");
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
    // default: level-synchronous, one warp per brick
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

int vrk_sdf_unbrick(vr_ctx* ctx, const int8_t* field, int nx, int ny, int nz, int8_t* linear) {
  BrickDims g{nx, ny, nz, nx / BR + 1, ny / BR + 1, nz / BR + 1};
  const size_t n = (size_t)nx * ny * nz;
  const unsigned blocks = (unsigned)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 16);
  k_sdf_unbrick<<<blocks, 256, 0, ctx->stream>>>(g, field, linear);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
