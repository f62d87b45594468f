// vr_sdf_common.cuh — definitions shared by the SDF build (vr_sdf.cu) and its alternative schedules (vr_sdf_variants.cu):
// the bricked field layout, the bit-volume geometry and its word-level dilation helpers, and the kernels both use
// (event bits, band bits, field assembly).  Kernels are `static`: each translation unit gets its own copy.
#pragma once
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

// ---- bit volumes: uint32 [z][y][ceil(nx/32)], one bit per voxel, 32 voxels along x per word -------------------------------------
#define WT_XW 4   // tile = 4 words (128 voxels) x 8 rows x 8 slices: one thread per word
#define WT_Y 8
#define WT_Z 8
#define WAVE_THREADS (WT_XW * WT_Y * WT_Z)

struct WaveDims {
  int nx, ny, nz;
  int nxw;         // words per row
  int bx, by, bz;  // bricks per axis of the field
  int tx, ty, tz;  // tiles per axis
  unsigned lastbit;  // bit of x == nx-1 in the last word of a row
};

__device__ __forceinline__ uint32_t shl_clamped(uint32_t c, uint32_t l, bool first) {  // bit i <- x-1 (x == 0 sees itself)
  return first ? ((c << 1) | (c & 1u)) : __funnelshift_l(l, c, 1);
}
__device__ __forceinline__ uint32_t shr_clamped(uint32_t c, uint32_t r, bool last, unsigned lastbit) {  // bit i <- x+1
  return last ? ((c >> 1) | (c & (1u << lastbit))) : __funnelshift_r(c, r, 1);
}
__device__ __forceinline__ uint32_t valid_mask(const WaveDims& g, int xw) {
  if (xw != g.nxw - 1 || g.lastbit == 31u) return 0xFFFFFFFFu;
  return (2u << g.lastbit) - 1u;
}

// E = event bit of every voxel.  Vector path (nx % 8 == 0, no gradient clause): a warp covers 256 voxels of a row with one
// 16-byte load per lane (512 contiguous bytes), each lane evaluates its 8 voxels, two shuffles assemble the words.
static __global__ void __launch_bounds__(256) k_sdf_events_v8(VolView vol, TfTable tf, int nxw, uint32_t* __restrict__ E,
                                                       unsigned chunks_per_row, unsigned nitems) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned it = warp; it < nitems; it += nwarps) {
    const unsigned row = it / chunks_per_row, chunk = it - row * chunks_per_row;
    const int x = (int)(chunk * 256 + lane * 8);
    unsigned bits = 0;
    if (x < vol.nx) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(vol.v + (size_t)row * vol.nx + x));
      const int w[4] = {q.x, q.y, q.z, q.w};
      if (tf.n == 1) {  // the two forms the UI / the tests generate most: one rectangle or one threshold (warp-uniform branch)
        const bool thr = tf.r[0].flags & VR_TF_THRESHOLD;
        const float lo = tf.r[0].min_v, hi = tf.r[0].max_v;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float fv = (float)(int)(short)((unsigned)w[k >> 1] >> (16 * (k & 1)));
          const bool e = thr ? fv > lo : (fv >= lo && fv <= hi);
          bits |= (e ? 1u : 0u) << k;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int v = (int)(short)((unsigned)w[k >> 1] >> (16 * (k & 1)));
          bits |= (tf_match(tf, v, 0) != 0 ? 1u : 0u) << k;
        }
      }
    }
    unsigned word = bits << (8 * (lane & 3));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    const int xw = (int)(chunk * 8 + (lane >> 2));
    if ((lane & 3) == 0 && xw < nxw) E[(size_t)row * nxw + xw] = word;
  }
}

// general path: any nx, TFs with a gradient clause (6 more taps per voxel through L1/L2)
template <bool GRAD>
static __global__ void __launch_bounds__(256) k_sdf_events(VolView vol, TfTable tf, int nxw, uint32_t* __restrict__ E,
                                                    unsigned nwords) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned w = warp; w < nwords; w += nwarps) {
    const unsigned row = w / (unsigned)nxw, xw = w - row * (unsigned)nxw;
    const int z = (int)(row / (unsigned)vol.ny), y = (int)(row - (unsigned)z * (unsigned)vol.ny);
    const int x = (int)(xw * 32 + lane);
    bool e = false;
    if (x < vol.nx) {
      if (GRAD) e = voxel_event(vol, tf, x, y, z) != 0;
      else e = tf_match(tf, __ldg(vol.v + (size_t)row * vol.nx + x), 0) != 0;
    }
    const unsigned bits = __ballot_sync(0xffffffffu, e);
    if (lane == 0) E[w] = bits;
  }
}

// spread the low 4 bits of b to 4 bytes 0x00/0x01
__device__ __forceinline__ uint32_t bits4_to_bytes(uint32_t b) { return ((b & 0xFu) * 0x00204081u) & 0x01010101u; }


// band bits only (R_0 into both bit volumes and into plane 0 = value 1); the field is written by k_sdf_assemble
static __global__ void __launch_bounds__(256) k_sdf_band_bits(WaveDims g, const uint32_t* __restrict__ E, uint32_t* __restrict__ Ra,
                                                       uint32_t* __restrict__ Rb, uint32_t* __restrict__ plane0, unsigned nwords) {
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
    plane0[w] = band;
  }
}

// planes + event bits -> bricked int8 field.  Same mapping as k_sdf_band: warp = word column x 8 rows of one z,
// lane = (8-bit piece) * 8 + row, so the 8 lanes of a piece write the 64 contiguous bytes of one z-slice of a brick.
static __global__ void __launch_bounds__(256) k_sdf_assemble(WaveDims g, int max_it, const uint32_t* __restrict__ E,
                                                      const uint32_t* __restrict__ planes, unsigned nwords,
                                                      int8_t* __restrict__ field, unsigned nxwf, unsigned items,
                                                             cudaSurfaceObject_t surf) {
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
      uint32_t m0 = 0, m1 = 0;  // per byte: the 7-bit level
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const uint32_t p8 = (__ldg(planes + (size_t)j * nwords + w) >> sh) & 0xFFu;
        m0 |= bits4_to_bytes(p8) << j;
        m1 |= bits4_to_bytes(p8 >> 4) << j;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t mag = h ? m1 : m0;
        // bytes that are 0 (never reached) become max_it: (mag | 0x80808080) - 0x01010101 has bit 7 clear exactly in zero bytes
        const uint32_t nz = (((mag | 0x80808080u) - 0x01010101u) >> 7) & 0x01010101u;  // 1 where the byte is non-zero
        mag |= (0x01010101u - nz) * (uint32_t)max_it;
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

