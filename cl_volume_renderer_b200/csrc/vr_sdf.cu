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
//   word) a level is a handful of shifts and ORs per word on a 16 MiB bit volume (512^3) that the previous level left in L2.
//     k_sdf_events    : E = event bit of every voxel (is_event_gen once per voxel; the reference evaluates it 9x)
//     k_sdf_band_bits9: R_0 = voxels with a clamped corner of the other event state (signed_distance_field.cl:22-48)
//     k_sdf_wave9     : one launch per level (programmatic dependent launch), R_{k-1} -> R_k, every R_k KEPT; a thread owns 128
//                       voxels of YR rows and walks TZ planes, nothing is recorded per voxel
//     k_sdf_count     : level of a voxel = 1 + the number of kept bit volumes in which its bit is still clear (bit-sliced counting)
//     k_sdf_assemble8 : counts + event bits -> the bricked int8 field (8x8x8 bricks of 512 bytes, apron of zeros at
//                       x == nx / y == ny / z == nz) and the 3-D array the ray marcher gathers from, written exactly once
//   Rows that are not a multiple of 4 words (and volumes too large to keep max_it - 1 bit volumes) use k_sdf_wave6: two bit
//   volumes, tile skipping, the level recorded at the moment a bit appears (RED.OR into 7 interleaved level planes).
//   The build is an object that advances level by level (vr_sdf_slab): the single-GPU build runs it to the end, the z-slab
//   sharded build (vr_comm.cu) swaps halo planes of the current bit volume between its ranks every K levels.
// Alternative schedules, all bit-exact, live in tools/ab/vr_sdf_variants.cu (and k_sdf_wave5 below) and are linked into the A/B
// build only (`make ab` -> libvr_ab.so, selected there with VR_SDF_MODE / VR_SDF_WAVE; DESIGN.md §4.2 has the table).
#include <cstring>
#include "vr_sdf_common.cuh"

// ---- one level: R_out = R_in | dilate(R_in) -------------------------------------------------------------------------------
// k_sdf_wave5 (round 1's level kernel, A/B build only): one WARP per tile (4 words x 8 rows x 8 planes), lane = lx + 4*ly, the
// warp walks the planes.  The y- and x-dilated rows yd(z') are computed once per plane (10 per tile) and reused by the planes z'-1 and z'+1, and all the
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

// ---- the level kernel for any row width (two bit volumes + level planes) -------------------------------------------------------
// k_sdf_wave5 above spends ~105 thread instructions per word and level (ncu: 13.9 M warp instructions per level at 512^3), most
// of them on per-lane halo handling: every lane loads the rows y-1 and y+1 itself and x-dilates both.  Here a THREAD owns a
// column of YR rows of one word and walks TZ planes: per plane it loads YR+2 words, x-dilates each once (the neighbouring words'
// edge bits come from the neighbouring lanes), ORs rows j-1 / j+1 in registers (no shuffles along y, no idle halo lanes) and
// keeps the previous two planes' results and its own words in registers, so a word is loaded once per tile and never re-read
// for the output: ~25 instructions per word and level.  A warp = XW words x GY row groups x GZ plane groups (one stamp tile).
// Level planes, interleaved: the 8 level words (7 used) of the voxel words w and w+1 (w even) share one 64-byte line,
// plane b of word w at (w >> 1) * 16 + 2 * b + (w & 1) — a pair's words of one plane are one 64-bit RED.OR target.
__device__ __forceinline__ size_t plane_word(unsigned w) { return (size_t)(w >> 1) * 16u + (w & 1u); }

template <int XW, int GY, int YR, int TZ, bool EDGE>  // EDGE: rows are wider than a tile (nxw > XW): edge lanes load their x neighbours
__device__ __forceinline__ void wave6_level(const WaveDims& g, int tx, int ty, int tz, int level, const uint32_t* Rin, uint32_t* Rout,
                                            uint32_t* planes, const int* stamp_in, int* stamp_out,
                                            unsigned* changed_tiles, bool all_active, int warp0, int nwarps) {
  constexpr int GZ = 32 / (XW * GY);
  static_assert(XW * GY * GZ == 32, "a warp is one tile");
  const int ntiles = tx * ty * tz;
  const unsigned lane = threadIdx.x & 31;
  const int lx = lane % XW, gy = (lane / XW) % GY, gz = lane / (XW * GY);
  const unsigned plane_stride = (unsigned)g.ny * (unsigned)g.nxw;
  const unsigned lv = (unsigned)level + 1u;
  // every read of the bit volumes and stamps goes to L2 (ld.cg): the persistent kernel reads what other SMs wrote one level ago
  for (int tile = warp0; tile < ntiles; tile += nwarps) {
    if (level != 1 && !all_active && __ldcg(stamp_in + tile) != level) continue;  // warp-uniform
    const int ttx = tile % tx, tq = tile / tx;
    const int tty = tq % ty, ttz = tq / ty;
    const int xw = ttx * XW + lx, y0 = (tty * GY + gy) * YR, z0 = (ttz * GZ + gz) * TZ;
    const bool xin = xw < g.nxw;
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const uint32_t lm = last ? (1u << g.lastbit) : 0u;  // x == nx-1 is its own +1 neighbour
    const uint32_t vm = valid_mask(g, xw);
    const bool lzero = lx == 0 || !xin, rzero = lx == XW - 1 || last || !xin;
    const bool lload = EDGE && lx == 0 && xw > 0 && xin, rload = EDGE && lx == XW - 1 && xw + 1 < g.nxw;
    unsigned ro[YR + 2];  // word offsets of the rows y0-1 .. y0+YR inside a plane, clamped like the corner coordinates
#pragma unroll
    for (int j = 0; j < YR + 2; ++j) ro[j] = (unsigned)min(max(y0 - 1 + j, 0), g.ny - 1) * (unsigned)g.nxw + (unsigned)min(xw, g.nxw - 1);
    uint32_t ydA[YR], ydB[YR], cprev[YR];  // x/y-dilated rows of the planes k-2 and k-1, the thread's own words of plane k-1
#pragma unroll
    for (int j = 0; j < YR; ++j) { ydA[j] = 0u; ydB[j] = 0u; cprev[j] = 0u; }
    bool changed = false;
    unsigned wout = (unsigned)z0 * plane_stride + (unsigned)y0 * (unsigned)g.nxw + (unsigned)xw;  // word of (row y0, plane z0 + k - 2)
    uint32_t c[YR + 2];  // the plane being processed; the next plane's words are requested before this one's are used
    {
      const unsigned zo = (unsigned)min(max(z0 - 1, 0), g.nz - 1) * plane_stride;  // 32-bit word indices: nwords <= 2^27
#pragma unroll
      for (int j = 0; j < YR + 2; ++j) c[j] = xin ? __ldcg(Rin + (zo + ro[j])) : 0u;
    }
#pragma unroll 1
    for (int k = 0; k < TZ + 2; ++k) {
      const unsigned zo = (unsigned)min(max(z0 - 1 + k, 0), g.nz - 1) * plane_stride;
      const unsigned zn = (unsigned)min(max(z0 + k, 0), g.nz - 1) * plane_stride;
      uint32_t cn[YR + 2], xd[YR + 2];
#pragma unroll
      for (int j = 0; j < YR + 2; ++j) cn[j] = (xin && k < TZ + 1) ? __ldcg(Rin + (zn + ro[j])) : 0u;
#pragma unroll
      for (int j = 0; j < YR + 2; ++j) {
        uint32_t l = __shfl_up_sync(0xffffffffu, c[j], 1), r = __shfl_down_sync(0xffffffffu, c[j], 1);
        if (lzero) l = 0u;
        if (rzero) r = 0u;
        if (EDGE) {
          if (lload) l = __ldcg(Rin + (zo + ro[j] - 1u));
          if (rload) r = __ldcg(Rin + (zo + ro[j] + 1u));
        }
        if (first) l = c[j] << 31;  // x == 0 is its own -1 neighbour
        xd[j] = __funnelshift_l(l, c[j], 1) | __funnelshift_r(c[j], r, 1) | (c[j] & lm);
      }
      // plane z0 + k - 2: its own words were loaded one iteration ago, its lower neighbour's rows two iterations ago
      if (k >= 2 && xin && z0 + k - 2 < g.nz) {
#pragma unroll
        for (int j = 0; j < YR; ++j) {
          if (y0 + j < g.ny) {
            const uint32_t old = cprev[j];
            const uint32_t now = (old | ydA[j] | xd[j] | xd[j + 2]) & vm;
            const uint32_t diff = now & ~old;
            const unsigned w = wout + (unsigned)j * (unsigned)g.nxw;
            Rout[w] = now;
            if (diff) {
              changed = true;
              uint32_t* pw = planes + plane_word(w);  // interleaved planes
#pragma unroll
              for (int b = 0; b < 7; ++b)
                if ((lv >> b) & 1u) atomicOr(pw + 2 * b, diff);  // result unused: RED.OR
            }
          }
        }
      }
      if (k >= 2) wout += plane_stride;
#pragma unroll
      for (int j = 0; j < YR; ++j) { ydA[j] = ydB[j]; ydB[j] = xd[j] | xd[j + 2]; cprev[j] = c[j + 1]; }
#pragma unroll
      for (int j = 0; j < YR + 2; ++j) c[j] = cn[j];
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

template <int XW, int GY, int YR, int TZ, bool EDGE>
__global__ void __launch_bounds__(128) k_sdf_wave6(WaveDims g, int tx, int ty, int tz, int level, const uint32_t* Rin, uint32_t* Rout,
                                                   uint32_t* planes, const int* stamp_in, int* stamp_out, unsigned* changed_tiles,
                                                   int all_active) {
  wave6_level<XW, GY, YR, TZ, EDGE>(g, tx, ty, tz, level, Rin, Rout, planes, stamp_in, stamp_out, changed_tiles, all_active != 0,
                                    (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), (int)((gridDim.x * blockDim.x) >> 5));
}

// ---- the default level kernel: 128 voxels of a row per thread, no level bits ------------------------------------------------------
// What the profiles of the kernels above said (512^3, ncu): the RED.ORs that record a voxel's level ARE the level's cost — the
// band of a CT is everywhere, so at mid build 80 % of all 128-voxel row pieces gain a bit in every level, about one bit each:
// 1.4-6.6 M RED sectors per level (2 x popcount(level+1) per piece), ~3.8 us per million, and the 134 MB of level planes push the
// two bit volumes out of L2 (45-60 MB of DRAM reads per level).  So this kernel records nothing: level k reads the bit volume R_{k-1}
// and writes R_k into its OWN buffer, all max_it-1 of them are kept (2 GiB at 512^3), and k_sdf_count recovers every voxel's level
// afterwards as the number of bit volumes in which its bit is still clear.  A level is then a pure streaming stencil
// (16 MiB in from L2, 16 MiB out), with no atomics, no tile stamps and no divergent tail.
//   A thread owns FOUR consecutive words (one 16-byte load) of YR rows: the x-dilation of the three inner word boundaries stays
// inside the thread (funnel shifts between its own registers), one shuffle pair per row quad fetches the outer two bits, and the
// rows above / below a thread's own come from the neighbouring lanes' registers (lane = XL quads along x times 32/XL row groups
// along y; the two rows outside the warp's range are x-dilated once per warp by its first / last row group); the warp walks TZ
// planes and keeps the previous two planes' results in registers.  Rows must be a multiple of 4 words (nx % 128 == 0, or
// nxw % 4 == 0 in general); other sizes use k_sdf_wave6 with its level planes.
//   flags[level] = 1 when the level set a bit, flags[130 + level] = 1 when it ran: once a level sets nothing the following ones
// return at once (their bit volumes are not written: k_sdf_count reuses the last written one) unless `force` (an imported halo).
struct Quad { uint32_t x, y, z, w; };
__device__ __forceinline__ Quad q_or(const Quad& a, const Quad& b) { return Quad{a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w}; }
__device__ __forceinline__ Quad q_load(const uint32_t* p, bool pred) {
  Quad q{0u, 0u, 0u, 0u};
  if (pred) { const uint4 v = __ldcg(reinterpret_cast<const uint4*>(p)); q.x = v.x; q.y = v.y; q.z = v.z; q.w = v.w; }
  return q;
}
template <int D>
__device__ __forceinline__ Quad q_shfl_up(const Quad& a) {
  return Quad{__shfl_up_sync(0xffffffffu, a.x, D), __shfl_up_sync(0xffffffffu, a.y, D), __shfl_up_sync(0xffffffffu, a.z, D),
              __shfl_up_sync(0xffffffffu, a.w, D)};
}
template <int D>
__device__ __forceinline__ Quad q_shfl_down(const Quad& a) {
  return Quad{__shfl_down_sync(0xffffffffu, a.x, D), __shfl_down_sync(0xffffffffu, a.y, D), __shfl_down_sync(0xffffffffu, a.z, D),
              __shfl_down_sync(0xffffffffu, a.w, D)};
}
struct XEdge {  // what a lane needs to close the x-dilation of its quad at the quad's two outer bits
  bool first, lzero, rzero, lload, rload;
  uint32_t lm;
};
template <bool EDGE>
__device__ __forceinline__ Quad x_dilate4(const Quad& a, const XEdge& e, const uint32_t* Rin, unsigned idx) {
  uint32_t l = __shfl_up_sync(0xffffffffu, a.w, 1), r = __shfl_down_sync(0xffffffffu, a.x, 1);
  if (e.lzero) l = 0u;
  if (e.rzero) r = 0u;
  if (EDGE) {
    if (e.lload) l = __ldcg(Rin + (idx - 1u));
    if (e.rload) r = __ldcg(Rin + (idx + 4u));
  }
  if (e.first) l = a.x << 31;  // x == 0 is its own -1 neighbour
  Quad d;
  d.x = __funnelshift_l(l, a.x, 1) | __funnelshift_r(a.x, a.y, 1);
  d.y = __funnelshift_l(a.x, a.y, 1) | __funnelshift_r(a.y, a.z, 1);
  d.z = __funnelshift_l(a.y, a.z, 1) | __funnelshift_r(a.z, a.w, 1);
  d.w = __funnelshift_l(a.z, a.w, 1) | __funnelshift_r(a.w, r, 1) | (a.w & e.lm);  // x == nx-1 is its own +1 neighbour
  return d;
}

#define SDF_FLAGS 130  // flags: changed[SDF_FLAGS] | ran[SDF_FLAGS]
// one warp tile of one level: R_out = R_in | dilate(R_in) on XL*4 words x (32/XL)*YR rows x TZ planes; returns "a bit was set"
template <int XL, int YR, int TZ, bool EDGE>
__device__ __forceinline__ bool wave9_tile(const WaveDims& g, int tx, int ty, int tile, const uint32_t* Rin, uint32_t* Rout) {
  constexpr int GYL = 32 / XL;
  static_assert(GYL >= 2 && XL * GYL == 32, "row groups");
  const unsigned lane = threadIdx.x & 31;
  const int lx = lane % XL, gy = lane / XL;
  const int nq = g.nxw >> 2;  // quads per row
  const unsigned plane_stride = (unsigned)g.ny * (unsigned)g.nxw;
  bool changed = false;
  {
    const int ttx = tile % tx, tq = tile / tx;
    const int tty = tq % ty, ttz = tq / ty;
    const int xq = ttx * XL + lx, y0 = (tty * GYL + gy) * YR, z0 = ttz * TZ;
    const bool xin = xq < nq, last = xq == nq - 1;
    XEdge e;
    e.first = xq == 0;
    e.lzero = lx == 0 || !xin;
    e.rzero = lx == XL - 1 || last || !xin;
    e.lload = EDGE && lx == 0 && xq > 0 && xin;
    e.rload = EDGE && lx == XL - 1 && xq + 1 < nq;
    e.lm = last ? (1u << g.lastbit) : 0u;
    const uint32_t vm3 = last ? valid_mask(g, g.nxw - 1) : 0xFFFFFFFFu;
    const unsigned xo = (unsigned)min(xq, nq - 1) * 4u;
    unsigned ro[YR];  // word offsets of the thread's rows inside a plane, clamped like the corner coordinates
#pragma unroll
    for (int j = 0; j < YR; ++j) ro[j] = (unsigned)min(y0 + j, g.ny - 1) * (unsigned)g.nxw + xo;
    // the row above the warp's range (first row group) or below it (last row group)
    const bool top = gy == 0, bot = gy == GYL - 1;
    const unsigned rh = (unsigned)min(max(top ? y0 - 1 : y0 + YR, 0), g.ny - 1) * (unsigned)g.nxw + xo;
    const bool hin = xin && (top || bot);
    Quad ydA[YR], ydB[YR], cprev[YR];  // x/y-dilated rows of the planes k-2 and k-1, the thread's own words of plane k-1
#pragma unroll
    for (int j = 0; j < YR; ++j) { ydA[j] = Quad{0u, 0u, 0u, 0u}; ydB[j] = ydA[j]; cprev[j] = ydA[j]; }
    unsigned wout = (unsigned)z0 * plane_stride + (unsigned)y0 * (unsigned)g.nxw + (unsigned)xq * 4u;  // (row y0, plane z0 + k - 2)
    // The plane being processed and a queue of PF planes already requested: an iteration is ~350 cycles of dependent work, an L2
    // round trip more than twice that, so a plane is requested PF + 1 iterations before it is used (with one plane ahead the
    // loop waited for every load: a lone warp took ~7 us for the 10 planes of a tile)
    constexpr int PF = 2;
    Quad c[YR], ch, q[PF][YR], qh[PF];
    auto plane_offset = [&](int k) { return (unsigned)min(max(z0 - 1 + k, 0), g.nz - 1) * plane_stride; };  // 32-bit word indices: nwords <= 2^27
    {
      const unsigned zo = plane_offset(0);
#pragma unroll
      for (int j = 0; j < YR; ++j) c[j] = q_load(Rin + (zo + ro[j]), xin);
      ch = q_load(Rin + (zo + rh), hin);
#pragma unroll
      for (int i = 0; i < PF; ++i) {
        const unsigned zi = plane_offset(1 + i);
#pragma unroll
        for (int j = 0; j < YR; ++j) q[i][j] = q_load(Rin + (zi + ro[j]), xin && 1 + i < TZ + 2);
        qh[i] = q_load(Rin + (zi + rh), hin && 1 + i < TZ + 2);
      }
    }
#pragma unroll 1
    for (int k = 0; k < TZ + 2; ++k) {
      const unsigned zo = plane_offset(k);
      const unsigned zn = plane_offset(k + PF + 1);
      Quad cn[YR], cnh, xd[YR];
#pragma unroll
      for (int j = 0; j < YR; ++j) cn[j] = q_load(Rin + (zn + ro[j]), xin && k + PF + 1 < TZ + 2);
      cnh = q_load(Rin + (zn + rh), hin && k + PF + 1 < TZ + 2);
#pragma unroll
      for (int j = 0; j < YR; ++j) xd[j] = x_dilate4<EDGE>(c[j], e, Rin, zo + ro[j]);
      const Quad xh = x_dilate4<EDGE>(ch, e, Rin, zo + rh);
      Quad up = q_shfl_up<XL>(xd[YR - 1]), dn = q_shfl_down<XL>(xd[0]);
      if (top) up = xh;
      if (bot) dn = xh;
      Quad yd[YR];
#pragma unroll
      for (int j = 0; j < YR; ++j) yd[j] = q_or(j == 0 ? up : xd[j - 1], j == YR - 1 ? dn : xd[j + 1]);
      // plane z0 + k - 2: its own words were loaded one iteration ago, its lower neighbour's rows two iterations ago
      if (k >= 2 && xin && z0 + k - 2 < g.nz) {
#pragma unroll
        for (int j = 0; j < YR; ++j) {
          if (y0 + j < g.ny) {
            const Quad old = cprev[j];
            Quad now = q_or(q_or(old, ydA[j]), yd[j]);
            now.w &= vm3;
            *reinterpret_cast<uint4*>(Rout + (wout + (unsigned)j * (unsigned)g.nxw)) = make_uint4(now.x, now.y, now.z, now.w);
            changed |= ((now.x ^ old.x) | (now.y ^ old.y) | (now.z ^ old.z) | (now.w ^ old.w)) != 0u;
          }
        }
      }
      if (k >= 2) wout += plane_stride;
#pragma unroll
      for (int j = 0; j < YR; ++j) {
        ydA[j] = ydB[j]; ydB[j] = yd[j]; cprev[j] = c[j]; c[j] = q[0][j];
#pragma unroll
        for (int i = 0; i + 1 < PF; ++i) q[i][j] = q[i + 1][j];
        q[PF - 1][j] = cn[j];
      }
      ch = qh[0];
#pragma unroll
      for (int i = 0; i + 1 < PF; ++i) qh[i] = qh[i + 1];
      qh[PF - 1] = cnh;
    }
    }
  return changed;
}

template <int XL, int YR, int TZ, bool EDGE, int MINB>
__global__ void __launch_bounds__(128, MINB) k_sdf_wave9(WaveDims g, int tx, int ty, int tz, int level, const uint32_t* Rin, uint32_t* Rout,
                                                         unsigned* flags, int force) {
  // programmatic dependent launch: the next level's CTAs may be scheduled as soon as this grid leaves room and do their index
  // arithmetic; they wait below until this grid has completed and its bit volume is visible
  cudaTriggerProgrammaticLaunchCompletion();
  const int ntiles = tx * ty * tz;
  const unsigned lane = threadIdx.x & 31;
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  bool changed = false;
  cudaGridDependencySynchronize();
  if (level > 1 && !force && __ldcg(flags + level - 1) == 0u) return;  // grid-uniform: the previous level set nothing
  if (blockIdx.x == 0 && threadIdx.x == 0) flags[SDF_FLAGS + level] = 1u;
  for (int tile = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); tile < ntiles; tile += nwarps)
    changed |= wave9_tile<XL, YR, TZ, EDGE>(g, tx, ty, tile, Rin, Rout);
  if (__any_sync(0xffffffffu, changed) && lane == 0) flags[level] = 1u;
}

// ---- all levels of a call in ONE launch, tiles synchronised point to point -----------------------------------------------------------
// A level per launch costs ~5.5 us of launch, ramp and drain whatever the volume's size (512^3: 11.3 us per level, a volume a fifth
// of it: 6.8 us), and a grid barrier costs no less.  But a tile of level k needs only the 3x3x3 tile neighbourhood of level k-1,
// and because every level writes its OWN bit volume there are no write-after-read hazards at all: a warp may run ahead of the
// rest of the grid as far as its neighbours allow.  So the grid is launched once (cooperatively: all CTAs resident, a warp per
// tile when they fit), every tile publishes the last level it completed (fence + release store), and a warp acquires its 27
// neighbours' counters before it starts a tile of the next level.  Stalled neighbours skew the wave locally instead of stopping it.
// flags[2 * SDF_FLAGS] = 1 if a wait timed out (the host reports it; nothing then waits any longer).
// Polling load: relaxed at gpu scope (served by L2).  An acquire would add an invalidation of the whole L1 (ncu: CCTL.IVALL, 16 % of
// the kernel's stall samples) that nothing here needs — every read of a bit volume is an ld.cg, which does not look at L1 — and
// the bit-volume loads of a tile are issued only after the counters it depends on have been seen (they are behind the branch).
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
template <int XL, int YR, int TZ, bool EDGE>
__global__ void __launch_bounds__(128) k_sdf_flow(WaveDims g, int tx, int ty, int tz, int level0, int nlevels, uint32_t* snaps, size_t nwords,
                                                  int* done, unsigned* flags) {
  const int ntiles = tx * ty * tz;
  const unsigned lane = threadIdx.x & 31;
  const int warp0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  unsigned* const err = flags + 2 * SDF_FLAGS;
  for (int it = level0; it < level0 + nlevels; ++it) {
    const uint32_t* Rin = snaps + (size_t)(it - 1) * nwords;
    uint32_t* Rout = snaps + (size_t)it * nwords;
    bool changed = false;
    if (warp0 == 0 && lane == 0) flags[SDF_FLAGS + it] = 1u;  // the level ran (k_sdf_count)
    for (int tile = warp0; tile < ntiles; tile += nwarps) {
      {  // acquire: the 3x3x3 tile neighbourhood has completed level it - 1
        const int ttx = tile % tx, tq = tile / tx;
        const int tty = tq % ty, ttz = tq / ty;
        const int ox = (int)lane % 3 - 1, oy = ((int)lane / 3) % 3 - 1, oz = (int)lane / 9 - 1;
        const int ax = ttx + ox, ay = tty + oy, az = ttz + oz;
        const bool need = lane < 27 && (unsigned)ax < (unsigned)tx && (unsigned)ay < (unsigned)ty && (unsigned)az < (unsigned)tz;
        const int* f = done + ((az * ty + ay) * tx + ax);
        long long t0 = 0;
        bool ok = !need || ld_acquire(f) >= it - 1;
        while (!__all_sync(0xffffffffu, ok)) {
          if (!t0) t0 = clock64();
          __nanosleep(64);
          if (!ok) ok = ld_acquire(f) >= it - 1;
          int bail = 0;
          if (lane == 0) {
            if (clock64() - t0 > 4000000000ll) *reinterpret_cast<volatile unsigned*>(err) = 1u;  // ~2 s: a neighbour is not coming
            bail = *reinterpret_cast<volatile unsigned*>(err) != 0u;
          }
          if (__shfl_sync(0xffffffffu, bail, 0)) return;  // a wait timed out somewhere: every warp gives up
        }
      }
      changed |= wave9_tile<XL, YR, TZ, EDGE>(g, tx, ty, tile, Rin, Rout);
      __syncwarp();                                // the lanes' stores happen before lane 0's release (one fence per warp:
      if (lane == 0) st_release(done + tile, it);  // with a __threadfence() per lane in front, ncu showed two ERRBARs, 24 % of the stalls)
    }
    if (__any_sync(0xffffffffu, changed) && lane == 0) flags[it] = 1u;
  }
}

// level of every voxel from the kept bit volumes R_0 .. R_{n-1}: z = the number of them in which its bit is still clear (a voxel
// that first appears in R_k has level k + 1; R_k of a level that did not run = the last written one).  Bit-sliced: a thread owns a
// word, compresses 7 volumes at a time with 4 full adders and adds the 3-bit result into 7 counter words; it stops at the first
// volume in which its word is full.  Output = the interleaved level planes k_sdf_assemble8 reads: 7 counter words + the word of
// the last volume (1 = reached; a voxel never reached gets max_it there).
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
  s = a ^ b ^ c;
  cy = (a & b) | (c & (a | b));
}
// add the number of set inputs among z[0..6] (bit-sliced) into the 7-bit counters cnt
__device__ __forceinline__ void count_add7(const uint32_t z[7], uint32_t cnt[7]) {
  uint32_t s0, c0, s1, c1, s2, c2, s3, c3;
  full_add(z[0], z[1], z[2], s0, c0);
  full_add(z[3], z[4], z[5], s1, c1);
  full_add(s0, s1, z[6], s2, c2);    // weight 1: s2
  full_add(c0, c1, c2, s3, c3);      // weight 2: s3, weight 4: c3
  uint32_t cy, t;
  t = cnt[0] & s2; cnt[0] ^= s2; cy = t;                                   // + s2
  full_add(cnt[1], s3, cy, t, cy); cnt[1] = t;                             // + 2 s3
  full_add(cnt[2], c3, cy, t, cy); cnt[2] = t;                             // + 4 c3
#pragma unroll
  for (int b = 3; b < 7; ++b) { t = cnt[b] & cy; cnt[b] ^= cy; cy = t; }
}
// A thread owns FOUR consecutive words (rows are a multiple of 4 words here): seven 16-byte loads in flight per thread — with one
// word per thread the pass reached 3.9 TB/s of the 2 GiB it streams.
static __global__ void __launch_bounds__(256) k_sdf_count(WaveDims g, const uint32_t* __restrict__ snaps, unsigned nwords, int nsnaps,
                                                         const unsigned* __restrict__ flags, uint32_t* __restrict__ planes8) {
  // the levels that ran are a prefix (a level returns early only when its predecessor set nothing, and so do all after it)
  __shared__ int nrun;
  if (threadIdx.x == 0) {
    int n = 1;
    while (n < nsnaps && flags[SDF_FLAGS + n] != 0u) ++n;
    nrun = n;
  }
  __syncthreads();
  const int n = nrun;
  const unsigned nquads = nwords >> 2;
  const unsigned qrow = (unsigned)g.nxw >> 2;
  for (unsigned qi = blockIdx.x * blockDim.x + threadIdx.x; qi < nquads; qi += gridDim.x * blockDim.x) {
    const unsigned w = qi * 4u;
    const uint32_t vm3 = (qi % qrow == qrow - 1) ? valid_mask(g, g.nxw - 1) : 0xFFFFFFFFu;  // only a row's last word can be partial
    uint32_t cnt[4][7];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int b = 0; b < 7; ++b) cnt[c][b] = 0u;
    uint4 cur = make_uint4(0u, 0u, 0u, 0u);
    const uint4* p = reinterpret_cast<const uint4*>(snaps + w);
    const size_t stride = (size_t)nwords >> 2;  // uint4 per bit volume
    for (int k0 = 0; k0 < n; k0 += 7, p += 7 * stride) {
      uint4 v[7];
      const int m = n - k0;  // volumes left (this group: min(m, 7)); beyond the end nothing is clear
#pragma unroll
      for (int i = 0; i < 7; ++i) v[i] = (i < m) ? __ldcs(p + (size_t)i * stride) : make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
      for (int i = 0; i < 7; ++i) if (i < m) cur = v[i];
      const bool full_first = (v[0].x & v[0].y & v[0].z & (v[0].w | ~vm3)) == 0xFFFFFFFFu;
      if (full_first) break;  // full already in the group's first volume: nothing more to count
      uint32_t z[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) z[i] = ~v[i].x;
      count_add7(z, cnt[0]);
#pragma unroll
      for (int i = 0; i < 7; ++i) z[i] = ~v[i].y;
      count_add7(z, cnt[1]);
#pragma unroll
      for (int i = 0; i < 7; ++i) z[i] = ~v[i].z;
      count_add7(z, cnt[2]);
#pragma unroll
      for (int i = 0; i < 7; ++i) z[i] = ~v[i].w & vm3;
      count_add7(z, cnt[3]);
      if ((v[6].x & v[6].y & v[6].z & (v[6].w | ~vm3)) == 0xFFFFFFFFu) break;
    }
    // the words w, w+1 share a 64-byte line of the interleaved planes, w+2, w+3 the next: {plane b of w, plane b of w+1} pairs
    uint4* q = reinterpret_cast<uint4*>(planes8 + plane_word(w));
    q[0] = make_uint4(cnt[0][0], cnt[1][0], cnt[0][1], cnt[1][1]);
    q[1] = make_uint4(cnt[0][2], cnt[1][2], cnt[0][3], cnt[1][3]);
    q[2] = make_uint4(cnt[0][4], cnt[1][4], cnt[0][5], cnt[1][5]);
    q[3] = make_uint4(cnt[0][6], cnt[1][6], cur.x, cur.y);
    q[4] = make_uint4(cnt[2][0], cnt[3][0], cnt[2][1], cnt[3][1]);
    q[5] = make_uint4(cnt[2][2], cnt[3][2], cnt[2][3], cnt[3][3]);
    q[6] = make_uint4(cnt[2][4], cnt[3][4], cnt[2][5], cnt[3][5]);
    q[7] = make_uint4(cnt[2][6], cnt[3][6], cur.z, cur.w & vm3);
  }
}

// band bits for the interleaved planes: R_0 into both bit volumes and into level plane 0, zeros into the other planes of the
// word (no memset of the planes)
static __global__ void __launch_bounds__(256) k_sdf_band_bits8(WaveDims g, const uint32_t* __restrict__ E, uint32_t* __restrict__ Ra,
                                                              uint32_t* __restrict__ Rb, uint32_t* __restrict__ planes8, unsigned nwords) {
  for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
    const unsigned row = w / (unsigned)g.nxw;
    const int xw = (int)(w - row * (unsigned)g.nxw);
    const int z = (int)(row / (unsigned)g.ny), y = (int)(row - (unsigned)z * (unsigned)g.ny);
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const uint32_t own = __ldg(E + w);
    uint32_t band = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = min(max(y + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(z + ((q & 2) ? 1 : -1), 0), g.nz - 1);
      const uint32_t* r = E + ((size_t)zz * g.ny + yy) * g.nxw;
      const uint32_t c = __ldg(r + xw);
      const uint32_t l = first ? 0u : __ldg(r + xw - 1), rr = last ? 0u : __ldg(r + xw + 1);
      band |= (shl_clamped(c, l, first) ^ own) | (shr_clamped(c, rr, last, g.lastbit) ^ own);
    }
    band &= valid_mask(g, xw);
    Ra[w] = band;
    Rb[w] = band;
    uint32_t* p = planes8 + plane_word(w);
    p[0] = band;
#pragma unroll
    for (int b = 1; b < 8; ++b) p[2 * b] = 0u;
  }
}

// interleaved planes + event bits -> bricked int8 field: k_sdf_assemble's mapping, the 7 level words of a voxel word from one
// 64-byte line
template <bool COUNT>  // COUNT: the planes hold k_sdf_count's clear counts z (level = z + 1) and the reached bits in slot 7
static __global__ void __launch_bounds__(256) k_sdf_assemble8(WaveDims g, int max_it, const uint32_t* __restrict__ E,
                                                             const uint32_t* __restrict__ planes8, int8_t* __restrict__ field,
                                                             unsigned nxwf, unsigned items, cudaSurfaceObject_t surf) {
  const unsigned lane = threadIdx.x & 31;
  const int yr = lane & 7, piece = lane >> 3;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned it = warp; it < items; it += nwarps) {
    const unsigned t = it / nxwf;
    const int xw = (int)(it - t * nxwf);
    const int z = (int)(t / (unsigned)g.by), yg = (int)(t - (unsigned)z * (unsigned)g.by);
    const int y = yg * 8 + yr;
    const int brick_x = xw * 4 + piece;
    if (brick_x >= g.bx) continue;
    uint32_t out[2] = {0u, 0u};
    if (xw < g.nxw && y < g.ny && z < g.nz) {
      const unsigned w = ((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw;
      const int sh = 8 * piece;
      const uint32_t e8 = (__ldg(E + w) >> sh) & 0xFFu, v8 = (valid_mask(g, xw) >> sh) & 0xFFu;
      const uint32_t* pp = planes8 + plane_word(w);
      uint32_t pl[7];
#pragma unroll
      for (int j = 0; j < 7; ++j) pl[j] = __ldg(pp + 2 * j);
      const uint32_t r8 = COUNT ? ((__ldg(pp + 14) >> sh) & 0xFFu) : 0u;
      uint32_t m0 = 0, m1 = 0;  // per byte: the 7-bit level
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const uint32_t p8 = (pl[j] >> sh) & 0xFFu;
        m0 |= bits4_to_bytes(p8) << j;
        m1 |= bits4_to_bytes(p8 >> 4) << j;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t mag = h ? m1 : m0;
        if (COUNT) {  // z <= 125: no carry between the bytes
          const uint32_t rb = bits4_to_bytes(r8 >> (4 * h)) * 0xFFu;
          mag = ((mag + 0x01010101u) & rb) | (((uint32_t)max_it * 0x01010101u) & ~rb);
        } else {
          // bytes that are 0 (never reached) become max_it: (mag | 0x80808080) - 0x01010101 has bit 7 clear exactly in zero bytes
          const uint32_t nz = (((mag | 0x80808080u) - 0x01010101u) >> 7) & 0x01010101u;  // 1 where the byte is non-zero
          mag |= (0x01010101u - nz) * (uint32_t)max_it;
        }
        const uint32_t ev = bits4_to_bytes(e8 >> (4 * h)), vd = bits4_to_bytes(v8 >> (4 * h));
        out[h] = ((mag ^ (ev * 0xFFu)) + ev) & (vd * 0xFFu);
      }
    }
    const size_t brick = ((size_t)(z >> 3) * g.by + yg) * g.bx + brick_x;
    *reinterpret_cast<uint2*>(field + brick * BRV + ((z & 7) << 6) + (yr << 3)) = make_uint2(out[0], out[1]);
    if (surf && y < g.ny && z < g.nz) {  // the same 8 voxels into the 3-D array the marcher gathers from (no apron there)
      const int x0 = brick_x * 8;
      if (x0 + 8 <= g.nx) surf3Dwrite(make_uint2(out[0], out[1]), surf, x0, y, z);
      else
        for (int k = 0; x0 + k < g.nx; ++k) surf3Dwrite((signed char)((out[k >> 2] >> (8 * (k & 3))) & 0xFFu), surf, x0 + k, y, z);
    }
  }
}

// k_sdf_count's counters -> the bricked field, a WORD per lane: a warp item is 4 words x 8 rows of one plane, lane = word * 8 + row,
// so a lane reads its word's eight plane entries once (k_sdf_assemble8 reads them in each of the four lanes that share a word:
// 28 instructions per voxel, issue-bound at 0.17 ms) and loops over the word's four 8-voxel pieces; the eight rows of a word
// still write 64 contiguous bytes of a brick slice per piece.
static __global__ void __launch_bounds__(256) k_sdf_assemble9(WaveDims g, int max_it, const uint32_t* __restrict__ E,
                                                             const uint32_t* __restrict__ planes8, int8_t* __restrict__ field,
                                                             unsigned nxg, unsigned items, cudaSurfaceObject_t surf) {
  const unsigned lane = threadIdx.x & 31;
  const int yr = lane & 7, wl = lane >> 3;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned it = warp; it < items; it += nwarps) {
    const unsigned t = it / nxg;
    const int xw = (int)(it - t * nxg) * 4 + wl;
    const int z = (int)(t / (unsigned)g.by), yg = (int)(t - (unsigned)z * (unsigned)g.by);
    const int y = yg * 8 + yr;
    if (xw * 4 >= g.bx) continue;
    const bool in = xw < g.nxw && y < g.ny && z < g.nz;
    uint32_t pl[7], ev = 0u, vm = 0u, reached = 0u;
#pragma unroll
    for (int j = 0; j < 7; ++j) pl[j] = 0u;
    if (in) {
      const unsigned w = ((unsigned)z * (unsigned)g.ny + (unsigned)y) * (unsigned)g.nxw + (unsigned)xw;
      ev = __ldg(E + w); vm = valid_mask(g, xw);
      const uint32_t* pp = planes8 + plane_word(w);
#pragma unroll
      for (int j = 0; j < 7; ++j) pl[j] = __ldg(pp + 2 * j);
      reached = __ldg(pp + 14);
    }
    const size_t slice = (((size_t)(z >> 3) * g.by + yg) * g.bx) * BRV + ((z & 7) << 6) + (yr << 3);
#pragma unroll
    for (int piece = 0; piece < 4; ++piece) {
      const int brick_x = xw * 4 + piece;
      if (brick_x >= g.bx) break;
      const int sh = 8 * piece;
      const uint32_t e8 = (ev >> sh) & 0xFFu, v8 = (vm >> sh) & 0xFFu, r8 = (reached >> sh) & 0xFFu;
      uint32_t m0 = 0, m1 = 0;  // per byte: the 7-bit count of bit volumes in which the voxel is clear
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const uint32_t p8 = (pl[j] >> sh) & 0xFFu;
        m0 |= bits4_to_bytes(p8) << j;
        m1 |= bits4_to_bytes(p8 >> 4) << j;
      }
      uint32_t out[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // level = count + 1 where the voxel was reached (count <= 125: no carry between the bytes), max_it elsewhere; sign by event
        const uint32_t rb = bits4_to_bytes(r8 >> (4 * h)) * 0xFFu;
        const uint32_t mag = (((h ? m1 : m0) + 0x01010101u) & rb) | (((uint32_t)max_it * 0x01010101u) & ~rb);
        const uint32_t evb = bits4_to_bytes(e8 >> (4 * h)), vd = bits4_to_bytes(v8 >> (4 * h));
        out[h] = ((mag ^ (evb * 0xFFu)) + evb) & (vd * 0xFFu);
      }
      *reinterpret_cast<uint2*>(field + slice + (size_t)brick_x * BRV) = make_uint2(out[0], out[1]);
      if (surf && in) {  // the same 8 voxels into the 3-D array the marcher gathers from (no apron there)
        const int x0 = brick_x * 8;
        if (x0 + 8 <= g.nx) surf3Dwrite(make_uint2(out[0], out[1]), surf, x0, y, z);
        else
          for (int k = 0; x0 + k < g.nx; ++k) surf3Dwrite((signed char)((out[k >> 2] >> (8 * (k & 3))) & 0xFFu), surf, x0 + k, y, z);
      }
    }
  }
}

// band bits only, into the first kept bit volume (k_sdf_wave9's scheme has no level planes to initialise)
static __global__ void __launch_bounds__(256) k_sdf_band_bits9(WaveDims g, const uint32_t* __restrict__ E, uint32_t* __restrict__ R0,
                                                              unsigned nwords) {
  for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
    const unsigned row = w / (unsigned)g.nxw;
    const int xw = (int)(w - row * (unsigned)g.nxw);
    const int z = (int)(row / (unsigned)g.ny), y = (int)(row - (unsigned)z * (unsigned)g.ny);
    const bool first = xw == 0, last = xw == g.nxw - 1;
    const uint32_t own = __ldg(E + w);
    uint32_t band = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = min(max(y + ((q & 1) ? 1 : -1), 0), g.ny - 1), zz = min(max(z + ((q & 2) ? 1 : -1), 0), g.nz - 1);
      const uint32_t* r = E + ((size_t)zz * g.ny + yy) * g.nxw;
      const uint32_t c = __ldg(r + xw);
      const uint32_t l = first ? 0u : __ldg(r + xw - 1), rr = last ? 0u : __ldg(r + xw + 1);
      band |= (shl_clamped(c, l, first) ^ own) | (shr_clamped(c, rr, last, g.lastbit) ^ own);
    }
    R0[w] = band & valid_mask(g, xw);
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
// level kernels: 9 = k_sdf_wave9 (rows of a multiple of 4 words; every level's bit volume kept, levels counted afterwards),
// 6 = k_sdf_wave6 (any size; two bit volumes + level planes), 5 = k_sdf_wave5 (A/B build; planar level planes)
#define W9_VARIANTS 6
#define W6_VARIANTS 3
struct vr_sdf_slab {
  vr_ctx* ctx = nullptr;
  WaveDims w{};
  int max_it = 0;      // of the GLOBAL volume (signed_distance_field.cpp:11)
  int level = 1;       // next level to run
  size_t nwords = 0, ntiles = 0;
  int wave = 9;        // level kernel
  int variant = -1;    // tile geometry of the kernel (-1: the kernel's default; A/B build: VR_SDF_VARIANT)
  int xl = 4;          // k_sdf_wave9: lanes along x
  int tile_z = WT_Z;   // planes per warp tile of k_sdf_wave5
  int nsnaps = 2;      // bit volumes behind E: max_it - 1 for k_sdf_wave9 (R_0 .. R_{max_it-2}), else the ping-pong pair
  uint32_t* scratch = nullptr;  // E | R[nsnaps] | stamps[2][ntiles] | flags: changed[130], ran[130]
  uint32_t* planes = nullptr;   // level planes (interleaved; planar for k_sdf_wave5)
  bool all_active = false;
  bool flow = true;         // k_sdf_wave9's tiles in one cooperative launch per call, synchronised point to point (k_sdf_flow)
  bool pdl = true;          // else: a launch per level with programmatic stream serialization
  bool early_exit = false;  // k_sdf_wave9: levels after one that set nothing return at once (single-GPU build only: a sharded build
                            // exchanges the current bit volume, which must then have been written)
  uint32_t* E() const { return scratch; }
  uint32_t* R(int i) const { return scratch + (size_t)(1 + i) * nwords; }
  int* stamps(int i) const { return reinterpret_cast<int*>(scratch + (size_t)(1 + nsnaps) * nwords + (size_t)i * ntiles); }
  unsigned* changed() const { return scratch + (size_t)(1 + nsnaps) * nwords + 2 * ntiles; }
  // the bit volume level `it` reads / writes
  uint32_t* Rin(int it) const { return wave == 9 ? R(it - 1) : R((it + 1) & 1); }
  uint32_t* Rout(int it) const { return wave == 9 ? R(it) : R(it & 1); }
};

// k_sdf_wave9 / k_sdf_flow tile shapes: {YR, TZ}.  The product uses 0 (k_sdf_flow) and 3 (k_sdf_wave9); measured at 512^3 and in a
// size sweep
static void w9_shape(int variant, int* yr, int* tz) {
  static const int t[W9_VARIANTS][2] = {{2, 8}, {2, 4}, {4, 4}, {4, 8}, {1, 8}, {2, 16}};
  *yr = t[variant][0]; *tz = t[variant][1];
}
// k_sdf_wave6 instantiations: {XW, GY, YR, TZ}
static void w6_shape(int variant, int* xw, int* gy, int* yr, int* tz) {
  static const int t[W6_VARIANTS][4] = {{16, 2, 4, 8}, {16, 2, 8, 8}, {16, 2, 4, 4}};
  *xw = t[variant][0]; *gy = t[variant][1]; *yr = t[variant][2]; *tz = t[variant][3];
}

int vrk_sdf_slab_create(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int max_it, vr_sdf_slab** out) {
  vr_sdf_slab* s = new (std::nothrow) vr_sdf_slab();
  if (!s) return VR_ERR_NOMEM;
  s->ctx = ctx; s->max_it = max_it;
  WaveDims& w = s->w;
  w.nx = nx; w.ny = ny; w.nz = nz;
  w.nxw = (nx + 31) / 32;
  w.bx = nx / BR + 1; w.by = ny / BR + 1; w.bz = nz / BR + 1;
  s->wave = (w.nxw % 4 == 0) ? 9 : 6;
#ifdef VR_AB
  // the A/B build can select the other level kernels and tile geometries (VR_SDF_WAVE=5|6|9, VR_SDF_VARIANT, VR_SDF_TZ for wave 5)
  static const int wave_env = getenv("VR_SDF_WAVE") ? atoi(getenv("VR_SDF_WAVE")) : 0;
  static const int var_env = getenv("VR_SDF_VARIANT") ? atoi(getenv("VR_SDF_VARIANT")) : 0;
  static const int tile_z_env = getenv("VR_SDF_TZ") ? atoi(getenv("VR_SDF_TZ")) : WT_Z;
  if (wave_env == 5 || wave_env == 6) s->wave = wave_env;
  s->variant = getenv("VR_SDF_VARIANT") ? std::max(var_env, 0) : -1;
  s->tile_z = (tile_z_env == 2 || tile_z_env == 4 || tile_z_env == 16) ? tile_z_env : WT_Z;
  static const int pdl_env = getenv("VR_SDF_PDL") ? atoi(getenv("VR_SDF_PDL")) : 1;
  static const int flow_env = getenv("VR_SDF_FLOW") ? atoi(getenv("VR_SDF_FLOW")) : -1;
  s->pdl = pdl_env != 0;
#endif
  w.lastbit = (unsigned)((nx - 1) & 31);
  s->nwords = (size_t)w.nxw * ny * nz;
  const size_t nwords = s->nwords;
  for (;;) {
    if (s->wave == 9) {
      // One launch per level (tiles of 4 rows x 8 planes per thread, programmatic dependent launch) while a level is a single
      // wave of CTAs at that kernel's 2 CTAs per SM; beyond that (640^3 and up on 148 SMs) the levels of a call run in one
      // cooperative launch whose tiles synchronise point to point (k_sdf_flow, 2 rows x 8 planes).  Measured, ms per build:
      // 256^3 0.95 / 1.06, 384^3 1.35 / 1.67, 512^3 1.79 / 2.12, 640^3 7.4 / 4.5, 768^3 11.2 / 7.1, 1024^3 34 / 15.0.
      s->xl = (w.nxw / 4 > 4) ? 8 : 4;
      const size_t tiles48 = (size_t)div_up(w.nxw / 4, s->xl) * div_up(ny, (32 / s->xl) * 4) * div_up(nz, 8);
      s->flow = tiles48 > (size_t)ctx->sm_count * 8;
#ifdef VR_AB
      if (flow_env >= 0) s->flow = flow_env != 0;
#endif
      if (s->variant < 0 || s->variant >= W9_VARIANTS) s->variant = s->flow ? 0 : 3;
      int yr, tz;
      w9_shape(s->variant, &yr, &tz);
      w.tx = div_up(w.nxw / 4, s->xl); w.ty = div_up(ny, (32 / s->xl) * yr); w.tz = div_up(nz, tz);
      s->nsnaps = std::max(max_it - 1, 1);
    } else if (s->wave == 6) {
      if (s->variant < 0 || s->variant >= W6_VARIANTS) s->variant = 0;
      int xw, gy, yr, tz;
      w6_shape(s->variant, &xw, &gy, &yr, &tz);
      w.tx = div_up(w.nxw, xw); w.ty = div_up(ny, gy * yr); w.tz = div_up(nz, (32 / (xw * gy)) * tz);
      s->nsnaps = 2;
    } else {
      w.tx = (w.nxw + WT_XW - 1) / WT_XW; w.ty = (ny + WT_Y - 1) / WT_Y; w.tz = (nz + s->tile_z - 1) / s->tile_z;
      s->nsnaps = 2;
    }
    s->ntiles = (size_t)w.tx * w.ty * w.tz;
    cudaError_t e = cudaMallocAsync(&s->scratch, ((size_t)(1 + s->nsnaps) * nwords + 2 * s->ntiles + 2 * SDF_FLAGS + 2) * 4, ctx->stream);
    if (e == cudaSuccess) {
      e = cudaMallocAsync(&s->planes, (s->wave == 5 ? 7 * nwords : 16 * ((nwords + 1) / 2)) * 4, ctx->stream);
      if (e != cudaSuccess) { cudaFreeAsync(s->scratch, ctx->stream); s->scratch = nullptr; }
    }
    if (e == cudaSuccess) break;
    cudaGetLastError();
    if (s->wave == 9) { s->wave = 6; s->variant = -1; continue; }  // no room for max_it - 1 bit volumes: two of them + level planes
    vr_set_error("vr_sdf_slab_create: %s", cudaGetErrorString(e));
    delete s;
    return VR_ERR_CUDA;
  }
  VR_CUDA(cudaMemsetAsync(s->stamps(0), 0, (2 * s->ntiles + 2 * SDF_FLAGS + 2) * 4, ctx->stream));
  if (s->wave == 5) VR_CUDA(cudaMemsetAsync(s->planes + nwords, 0, 6 * nwords * 4, ctx->stream));
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
  if (s->wave == 9) k_sdf_band_bits9<<<bb, 256, 0, ctx->stream>>>(w, s->E(), s->R(0), (unsigned)nwords);
  else if (s->wave == 6) k_sdf_band_bits8<<<bb, 256, 0, ctx->stream>>>(w, s->E(), s->R(0), s->R(1), s->planes, (unsigned)nwords);
  else k_sdf_band_bits<<<bb, 256, 0, ctx->stream>>>(w, s->E(), s->R(0), s->R(1), s->planes, (unsigned)nwords);
  ctx->launches += 2;
  VR_CUDA(cudaGetLastError());
  *out = s;
  return VR_OK;
}

template <int XL, int YR, int TZ, int MINB>
static void launch_wave9(vr_sdf_slab* s, int it, unsigned grid) {
  const WaveDims& w = s->w;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = s->ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = s->pdl ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const uint32_t* rin = s->Rin(it);
  const int force = (s->all_active || !s->early_exit) ? 1 : 0;
  if (w.nxw / 4 > XL) cudaLaunchKernelEx(&cfg, k_sdf_wave9<XL, YR, TZ, true, MINB>, w, w.tx, w.ty, w.tz, it, rin, s->Rout(it), s->changed(), force);
  else cudaLaunchKernelEx(&cfg, k_sdf_wave9<XL, YR, TZ, false, MINB>, w, w.tx, w.ty, w.tz, it, rin, s->Rout(it), s->changed(), force);
}
// cooperative launch of k_sdf_flow: the grid must be resident (occupancy x SMs, queried once per instantiation)
template <int XL, int YR, int TZ>
static int launch_flow(vr_sdf_slab* s, int level0, int nlevels) {
  vr_ctx* ctx = s->ctx;
  WaveDims w = s->w;
  const bool edge = w.nxw / 4 > XL;
  void (*kern)(WaveDims, int, int, int, int, int, uint32_t*, size_t, int*, unsigned*) =
      edge ? k_sdf_flow<XL, YR, TZ, true> : k_sdf_flow<XL, YR, TZ, false>;
  static int occ[2] = {0, 0};
  if (!occ[edge]) {
    VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[edge], kern, 128, 0));
    if (occ[edge] < 1) { vr_set_error("k_sdf_flow: no resident block"); return VR_ERR_CUDA; }
  }
  const unsigned grid = (unsigned)std::min<size_t>(div_up(s->ntiles, 4), (size_t)ctx->sm_count * occ[edge]);
  int tx = w.tx, ty = w.ty, tz = w.tz;
  uint32_t* snaps = s->R(0);
  size_t nwords = s->nwords;
  int* done = s->stamps(0);
  unsigned* flags = s->changed();
  void* args[] = {&w, &tx, &ty, &tz, &level0, &nlevels, &snaps, &nwords, &done, &flags};
  VR_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(128), args, 0, ctx->stream));
  return VR_OK;
}

template <int XW, int GY, int YR, int TZ>
static void launch_wave6(vr_sdf_slab* s, int it, unsigned grid) {
  const WaveDims& w = s->w;
  if (w.nxw > XW)
    k_sdf_wave6<XW, GY, YR, TZ, true><<<grid, 128, 0, s->ctx->stream>>>(w, w.tx, w.ty, w.tz, it, s->Rin(it), s->Rout(it), s->planes,
                                                                        s->stamps(it & 1), s->stamps((it + 1) & 1), s->changed(),
                                                                        s->all_active ? 1 : 0);
  else
    k_sdf_wave6<XW, GY, YR, TZ, false><<<grid, 128, 0, s->ctx->stream>>>(w, w.tx, w.ty, w.tz, it, s->Rin(it), s->Rout(it), s->planes,
                                                                         s->stamps(it & 1), s->stamps((it + 1) & 1), s->changed(),
                                                                         s->all_active ? 1 : 0);
}

int vrk_sdf_slab_advance(vr_sdf_slab* s, int nlevels, int* done) {
  vr_ctx* ctx = s->ctx;
  const unsigned grid = (unsigned)std::min<size_t>(div_up(s->ntiles, 4), (size_t)ctx->sm_count * 64);  // a warp per tile, CTAs of 4 warps
  int n = 0;
  if (s->wave == 9 && s->flow) {
    n = std::max(0, std::min(nlevels, s->max_it - 1 - s->level));
    if (n > 0) {
      int st;
#define VR_FLOW(YR, TZ) st = s->xl == 8 ? launch_flow<8, YR, TZ>(s, s->level, n) : launch_flow<4, YR, TZ>(s, s->level, n)
      switch (s->variant) {
#ifdef VR_AB
        case 1: VR_FLOW(2, 4); break;
        case 2: VR_FLOW(4, 4); break;
        case 3: VR_FLOW(4, 8); break;
        case 4: VR_FLOW(1, 8); break;
        case 5: VR_FLOW(2, 16); break;
#endif
        default: VR_FLOW(2, 8); break;  // variant 0
      }
#undef VR_FLOW
      if (st == VR_OK) {
        s->level += n;
        s->all_active = false;
        ctx->launches++;
        if (done) *done = n;
        return VR_OK;
      }
      // the cooperative launch was refused (no co-residency on this device / under this sharing mode): a launch per level with
      // the same tiles from here on
      cudaGetLastError();
      s->flow = false;
      n = 0;
    } else {
      if (done) *done = 0;
      return VR_OK;
    }
  }
  for (; n < nlevels && s->level + 1 < s->max_it; ++n, ++s->level) {
    const int it = s->level;
    if (s->wave == 9) {
#define VR_W9(YR, TZ, MINB) do { if (s->xl == 8) launch_wave9<8, YR, TZ, MINB>(s, it, grid); else launch_wave9<4, YR, TZ, MINB>(s, it, grid); } while (0)
      switch (s->variant) {
        case 0: VR_W9(2, 8, 4); break;  // k_sdf_flow's tile shape: the fallback when a cooperative launch is refused
#ifdef VR_AB
        case 1: VR_W9(2, 4, 4); break;
        case 2: VR_W9(4, 4, 2); break;
        case 4: VR_W9(1, 8, 6); break;
        case 5: VR_W9(2, 16, 4); break;
#endif
        default: VR_W9(4, 8, 2); break;  // variant 3
      }
#undef VR_W9
    } else if (s->wave == 6) {
      switch (s->variant) {
#ifdef VR_AB
        case 1: launch_wave6<16, 2, 8, 8>(s, it, grid); break;
        case 2: launch_wave6<16, 2, 4, 4>(s, it, grid); break;
#endif
        default: launch_wave6<16, 2, 4, 8>(s, it, grid); break;
      }
    } else {
#ifdef VR_AB
      const WaveDims& w = s->w;
      // measured at 512^3 (grid multiplier): 8..12 3.9 ms, 16 3.53, 32 (a warp per tile, no loop) 3.43
      static const int grid_mult = getenv("VR_SDF_GRID") ? std::max(atoi(getenv("VR_SDF_GRID")), 1) : 64;
      static const int cta_warps = getenv("VR_SDF_WARPS") ? std::min(std::max(atoi(getenv("VR_SDF_WARPS")), 1), 8) : 4;
      const unsigned wg5 = (unsigned)std::min<size_t>(div_up(s->ntiles, cta_warps), (size_t)ctx->sm_count * grid_mult * 4 / cta_warps);
#define VR_WAVE5(TZ)                                                                                                          \
  k_sdf_wave5<4, TZ><<<wg5, 32 * cta_warps, 0, ctx->stream>>>(w, w.tx, w.ty, w.tz, it, s->Rin(it), s->Rout(it), s->planes,        \
                                                              (unsigned)s->nwords, s->stamps(it & 1), s->stamps((it + 1) & 1),   \
                                                              s->changed(), s->all_active ? 1 : 0)
      switch (s->tile_z) {
        case 2: VR_WAVE5(2); break;
        case 4: VR_WAVE5(4); break;
        case 16: VR_WAVE5(16); break;
        default: VR_WAVE5(8); break;
      }
#undef VR_WAVE5
#else
      vr_set_error("vrk_sdf_slab_advance: unknown level kernel");
      return VR_ERR_INVALID;
#endif
    }
    s->all_active = false;
    ctx->launches++;
  }
  VR_CUDA(cudaGetLastError());
  if (done) *done = n;
  return VR_OK;
}

uint32_t* vrk_sdf_slab_bits(vr_sdf_slab* s) { return s->Rin(s->level); }  // the volume level `s->level` will read
size_t vrk_sdf_slab_plane_words(const vr_sdf_slab* s) { return (size_t)s->w.nxw * s->w.ny; }
void vrk_sdf_slab_mark_imported(vr_sdf_slab* s) { s->all_active = true; }
int vrk_sdf_slab_level(const vr_sdf_slab* s) { return s->level; }
bool vrk_sdf_slab_finished(const vr_sdf_slab* s) { return s->level + 1 >= s->max_it; }

int vrk_sdf_slab_assemble(vr_sdf_slab* s, int8_t* field, cudaSurfaceObject_t surf) {
  const WaveDims& w = s->w;
  const unsigned nxwf = (unsigned)((8 * w.bx + 31) / 32);
  const unsigned items = nxwf * (unsigned)w.by * (8u * (unsigned)w.bz);
  const unsigned bg = (unsigned)std::min<size_t>(div_up(items, 8), (size_t)s->ctx->sm_count * 16);
  if (s->wave == 9) {
    // the levels run so far: R_0 .. R_{level-1}
    const unsigned cg = (unsigned)std::min<size_t>(div_up(s->nwords / 4, 256), (size_t)s->ctx->sm_count * 16);
    k_sdf_count<<<cg, 256, 0, s->ctx->stream>>>(w, s->R(0), (unsigned)s->nwords, std::min(s->level, s->nsnaps), s->changed(), s->planes);
    s->ctx->launches++;
    const unsigned nxg = (nxwf + 3) / 4, items9 = nxg * (unsigned)w.by * (8u * (unsigned)w.bz);
    k_sdf_assemble9<<<(unsigned)std::min<size_t>(div_up(items9, 8), (size_t)s->ctx->sm_count * 16), 256, 0, s->ctx->stream>>>(
        w, s->max_it, s->E(), s->planes, field, nxg, items9, surf);
  } else if (s->wave == 6) {
    k_sdf_assemble8<false><<<bg, 256, 0, s->ctx->stream>>>(w, s->max_it, s->E(), s->planes, field, nxwf, items, surf);
  } else {
    k_sdf_assemble<<<bg, 256, 0, s->ctx->stream>>>(w, s->max_it, s->E(), s->planes, (unsigned)s->nwords, field, nxwf, items, surf);
  }
  s->ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// blocking: has a tile of k_sdf_flow given up waiting for a neighbour?  (never observed; the field would be incomplete)
int vrk_sdf_slab_status(vr_sdf_slab* s) {
  if (!(s->wave == 9 && s->flow)) return VR_OK;
  unsigned* pin = reinterpret_cast<unsigned*>(s->ctx->scratch_host);
  VR_CUDA(cudaMemcpyAsync(pin, s->changed() + 2 * SDF_FLAGS, sizeof(unsigned), cudaMemcpyDeviceToHost, s->ctx->stream));
  VR_CUDA(cudaStreamSynchronize(s->ctx->stream));
  if (pin[0] != 0u) { vr_set_error("vr_sdf: a tile of the level wave waited in vain for a neighbour"); return VR_ERR_CUDA; }
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

// Is a build of this size in the regime where a level is a single wave of CTAs (vrk_sdf_slab_create's criterion)?  There a level
// costs the serial chain of one warp through its tile whatever the number of planes, so z-slabs on several GPUs buy nothing
// (512^3: 1.61 ms on one GPU, 1.73 ms on 8) and only add the halo swaps and the gather; above it they pay (1024^3: 13.6 -> 5.7 ms).
bool vrk_sdf_single_wave(const vr_ctx* ctx, int nx, int ny, int nz) {
  const int nxw = (nx + 31) / 32;
  if (nxw % 4 != 0) return (size_t)nx * ny * nz <= ((size_t)1 << 27);  // the two-volume kernel: same order of magnitude
  const int xl = (nxw / 4 > 4) ? 8 : 4;
  const size_t tiles48 = (size_t)div_up(nxw / 4, xl) * div_up(ny, (32 / xl) * 4) * div_up(nz, 8);
  return tiles48 <= (size_t)ctx->sm_count * 8;
}

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
  s->early_exit = true;
  int st = vrk_sdf_slab_advance(s, max_it, nullptr);
  if (st == VR_OK) st = vrk_sdf_slab_assemble(s, field, surf);
  // diagnostics: the last level that set a bit
  int levels = 0;
  if (st == VR_OK) {
    unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
    cudaError_t e = cudaMemcpyAsync(hc, s->changed(), sizeof(unsigned) * (2 * SDF_FLAGS + 1), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { vr_set_error("vrk_sdf_build: %s", cudaGetErrorString(e)); st = VR_ERR_CUDA; }
    if (st == VR_OK && hc[2 * SDF_FLAGS] != 0) { vr_set_error("vrk_sdf_build: a tile of the level wave waited in vain for a neighbour"); st = VR_ERR_CUDA; }
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
