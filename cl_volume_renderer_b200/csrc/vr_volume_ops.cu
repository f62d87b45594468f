// vr_volume_ops.cu — streaming kernels over the whole volume: fetch_stats, apply_clip, bilateral_filter,
// tf_sort_values, tf_flush_color_frame, buffer_reset.  All are HBM- (or, for the bilateral filter, SFU-) bound;
// none is a contraction, so no tensor cores.
#include "vr_device.cuh"

// 3-D tile of a block: 32 voxels along x (one warp = one 64-byte row segment), 4 rows, 4 slices.  The y±1 and
// z±1 gradient taps of a voxel are the centre taps of other warps of the same block, so they hit in L1.
#define TX 32
#define TY 4
#define TZ 4

// ---- fetch_stats: reference_volume_figures.cl:10-26 -----------------------------------------------------------
// The reference issues four global atomics per voxel; here: per-thread -> warp shuffle -> block -> 4 atomics/block.
__global__ void __launch_bounds__(TX* TY* TZ) k_fetch_stats(VolView vol, int32_t* __restrict__ stats, int zlo, int zhi) {
  const int x = blockIdx.x * TX + threadIdx.x;
  const int y = blockIdx.y * TY + threadIdx.y;
  const int z = blockIdx.z * TZ + threadIdx.z;
  int mnv = INT32_MAX, mxv = INT32_MIN, mng = INT32_MAX, mxg = INT32_MIN;
  if (x < vol.nx && y < vol.ny && z >= zlo && z < zhi) {
    int v = vol.at(x, y, z);
    int g = f2i(length3(gradient_voxel(vol, x, y, z)));  // implicit float->int of atomic_min/max(int*, float)
    mnv = mxv = v;
    mng = mxg = g;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnv = min(mnv, __shfl_xor_sync(0xffffffffu, mnv, o));
    mxv = max(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
    mng = min(mng, __shfl_xor_sync(0xffffffffu, mng, o));
    mxg = max(mxg, __shfl_xor_sync(0xffffffffu, mxg, o));
  }
  __shared__ int s[4][TY * TZ];
  const int warp = threadIdx.y + TY * threadIdx.z;
  if (threadIdx.x == 0) {
    s[0][warp] = mnv; s[1][warp] = mxv; s[2][warp] = mng; s[3][warp] = mxg;
  }
  __syncthreads();
  if (warp == 0 && threadIdx.x < TY * TZ) {
    mnv = s[0][threadIdx.x]; mxv = s[1][threadIdx.x]; mng = s[2][threadIdx.x]; mxg = s[3][threadIdx.x];
    for (int o = TY * TZ / 2; o > 0; o >>= 1) {
      mnv = min(mnv, __shfl_xor_sync(0x0000ffffu, mnv, o));
      mxv = max(mxv, __shfl_xor_sync(0x0000ffffu, mxv, o));
      mng = min(mng, __shfl_xor_sync(0x0000ffffu, mng, o));
      mxg = max(mxg, __shfl_xor_sync(0x0000ffffu, mxg, o));
    }
    if (threadIdx.x == 0) {
      atomicMin(stats + 0, mnv); atomicMax(stats + 1, mxv);
      atomicMin(stats + 2, mng); atomicMax(stats + 3, mxg);
    }
  }
}

// ---- vectorised stencil front end for fetch_stats / tf_sort_values (nx % 8 == 0) ------------------------------------------
// A thread owns 8 consecutive voxels of a row: the centre row and its y+-1 / z+-1 rows are five 16-byte loads, the two x
// neighbours beyond the octet two scalar loads — 7 load instructions per 8 voxels instead of 56.  Rows that fall outside
// the volume read the border colour 0 (CLK_ADDRESS_CLAMP).
__device__ __forceinline__ unsigned div_up_dev(int a, int b) { return (unsigned)((a + b - 1) / b); }
struct Octet {
  int c[8];                // centre values
  int dx[8], dy[8], dz[8]; // central differences (not halved), utility_filter.cl:2-35
};
__device__ __forceinline__ void unpack8(const uint4 q, int v[8]) {
  v[0] = (short)(q.x & 0xFFFF); v[1] = (short)(q.x >> 16); v[2] = (short)(q.y & 0xFFFF); v[3] = (short)(q.y >> 16);
  v[4] = (short)(q.z & 0xFFFF); v[5] = (short)(q.z >> 16); v[6] = (short)(q.w & 0xFFFF); v[7] = (short)(q.w >> 16);
}
__device__ __forceinline__ Octet load_octet(const VolView& vol, int x0, int y, int z) {
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const size_t row = ((size_t)z * vol.ny + y) * vol.nx + x0;
  const size_t sy = vol.nx, sz = (size_t)vol.nx * vol.ny;
  const uint4* base = reinterpret_cast<const uint4*>(vol.v + row);
  const uint4 qc = __ldg(base);
  const uint4 qym = y > 0 ? __ldg(reinterpret_cast<const uint4*>(vol.v + row - sy)) : zero;
  const uint4 qyp = y + 1 < vol.ny ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sy)) : zero;
  const uint4 qzm = z > 0 ? __ldg(reinterpret_cast<const uint4*>(vol.v + row - sz)) : zero;
  const uint4 qzp = z + 1 < vol.nz ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sz)) : zero;
  const int xl = x0 > 0 ? (int)__ldg(vol.v + row - 1) : 0;
  const int xr = x0 + 8 < vol.nx ? (int)__ldg(vol.v + row + 8) : 0;
  Octet o;
  int ym[8], yp[8], zm[8], zp[8];
  unpack8(qc, o.c); unpack8(qym, ym); unpack8(qyp, yp); unpack8(qzm, zm); unpack8(qzp, zp);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int l = k == 0 ? xl : o.c[k - 1], r = k == 7 ? xr : o.c[k + 1];
    o.dx[k] = r - l; o.dy[k] = yp[k] - ym[k]; o.dz[k] = zp[k] - zm[k];
  }
  return o;
}
// squared gradient length in the evaluation order of length(): (dx*dx + dy*dy) + dz*dz in fp32
__device__ __forceinline__ float grad_sq(const Octet& o, int k) {
  const float fx = (float)o.dx[k], fy = (float)o.dy[k], fz = (float)o.dz[k];
  return (fx * fx + fy * fy) + fz * fz;
}

#define VX 16  // threads along x (128 voxels)
#define VY 8
#define VZ 2

// ---- VR_SAMPLING_HW_LINEAR for the volume kernels: the box-averaged volume ----------------------------------------------------------
// fetch_stats, tf_sort_values and bilateral_filter read the volume through CLK_FILTER_LINEAR samplers at INTEGER coordinates
// (reference_volume_figures.cl:14-23, histogram.cl:10-15, utility_filter.cl:38-62).  With texel centres at +0.5 an integer
// coordinate lies exactly between two texels on every axis: the hardware's fixed-point fractions are 128/256 and its eight
// weights come out as 32/256 each (oracle.cpp hw_linear_fetch: U = 64 four times, then 32 + 32), so the filtered value is
//     B(x,y,z) = floor( (sum of the 2x2x2 texels (x-1..x, y-1..y, z-1..z)) / 8 + 1/2 ) = (S + 4) >> 3
// — no texture unit needed.  k_boxavg writes B with the border addressing of gradient_prewitt_nn / bilateral_kernel (texels
// outside read 0) for coordinates 0..n inclusive, rows padded to a multiple of 8 with zeros: a volume of (pad8(nx+1), ny+1,
// nz+1) voxels whose out-of-range reads are 0 is exactly B everywhere, so the vectorised kernels below run on it unchanged.
// The VALUE read of fetch_stats / tf_sort_values uses a sampler without an addressing mode, served like clamp-to-edge: it
// differs from B only on the faces x == 0, y == 0, z == 0, where it is recomputed from the volume (box_edge).
__device__ __forceinline__ int box_border(const VolView& v, int x, int y, int z) {
  int s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) s += v.at(x - 1 + (c & 1), y - 1 + ((c >> 1) & 1), z - 1 + (c >> 2));
  return (s + 4) >> 3;
}
__device__ __forceinline__ int box_edge(const VolView& v, int x, int y, int z) {
  int s = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    s += v.at(min(max(x - 1 + (c & 1), 0), v.nx - 1), min(max(y - 1 + ((c >> 1) & 1), 0), v.ny - 1), min(max(z - 1 + (c >> 2), 0), v.nz - 1));
  return (s + 4) >> 3;
}
// A thread writes 8 consecutive x of one (y, z): the four contributing rows come as 16-byte loads (plus the voxel left of the group),
// summed per column first, then neighbouring columns — 4 vector loads and ~100 instructions per 8 outputs (first version: 8 scalar
// loads per output, 0.69 ms at 512^3).
__global__ void __launch_bounds__(256) k_boxavg(VolView v, int16_t* __restrict__ out, int px) {
  const unsigned gx = (unsigned)px >> 3;  // px is a multiple of 8
  const size_t ngroups = (size_t)gx * (v.ny + 1) * (v.nz + 1);
  const bool vec = (v.nx & 7) == 0;  // rows of the volume 16-byte aligned
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += (size_t)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % gx) * 8;
    const size_t t = i / gx;
    const int y = (int)(t % (v.ny + 1)), z = (int)(t / (v.ny + 1));
    int col[9];  // sums over the 2 x 2 rows (y-1..y, z-1..z) of the voxels x0-1 .. x0+7; voxels outside the volume are 0
#pragma unroll
    for (int k = 0; k < 9; ++k) col[k] = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int yy = y - 1 + (c & 1), zz = z - 1 + (c >> 1);
      if ((unsigned)yy >= (unsigned)v.ny || (unsigned)zz >= (unsigned)v.nz) continue;
      const int16_t* row = v.v + ((size_t)zz * v.ny + yy) * v.nx;
      if (x0 > 0 && x0 - 1 < v.nx) col[0] += (int)__ldg(row + x0 - 1);
      if (vec && x0 + 8 <= v.nx) {
        int w[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(row + x0)), w);
#pragma unroll
        for (int k = 0; k < 8; ++k) col[1 + k] += w[k];
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (x0 + k < v.nx) col[1 + k] += (int)__ldg(row + x0 + k);
      }
    }
    unsigned o[4];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const int a = x0 + k <= v.nx ? (col[k] + col[k + 1] + 4) >> 3 : 0, b = x0 + k + 1 <= v.nx ? (col[k + 1] + col[k + 2] + 4) >> 3 : 0;
      o[k >> 1] = ((unsigned)a & 0xFFFFu) | ((unsigned)b << 16);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
int vrk_boxavg(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int16_t* out, int px) {
  VolView v{vol, nx, ny, nz};
  const size_t n = (size_t)(px / 8) * (ny + 1) * (nz + 1);
  k_boxavg<<<(unsigned)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 32), 256, 0, ctx->stream>>>(v, out, px);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// fetch_stats, vectorised.  sqrt and float->int are monotonic, so min/max of (int)sqrt(s) = (int)sqrt(min/max s): the
// square root is taken once per thread instead of once per voxel — bit-identical.
// LINEAR: `vol` is the box-averaged volume, `orig` the volume itself; only voxels x < lim_x, y < lim_y count.
template <bool LINEAR>
__global__ void __launch_bounds__(VX* VY* VZ) k_fetch_stats_v8(VolView vol, VolView orig, int lim_x, int lim_y, int32_t* __restrict__ stats,
                                                               int zlo, int zhi) {
  int mnv = INT32_MAX, mxv = INT32_MIN, mng = INT32_MAX, mxg = INT32_MIN;
  const unsigned tx = div_up_dev(lim_x, VX * 8), ty = div_up_dev(lim_y, VY), tz = div_up_dev(zhi, VZ);
  for (unsigned t = blockIdx.x; t < tx * ty * tz; t += gridDim.x) {  // persistent CTAs: one set of atomics per CTA at the end
    const int x0 = (int)(((t % tx) * VX + threadIdx.x) * 8);
    const int y = (int)(((t / tx) % ty) * VY + threadIdx.y);
    const int z = (int)((t / (tx * ty)) * VZ + threadIdx.z);
    if (x0 < lim_x && y < lim_y && z >= zlo && z < zhi) {
      const Octet o = load_octet(vol, x0, y, z);
      float smin = grad_sq(o, 0), smax = smin;
      const bool face = LINEAR && (y == 0 || z == 0);
      const int v0 = (LINEAR && (face || x0 == 0)) ? box_edge(orig, x0, y, z) : o.c[0];
      mnv = min(mnv, v0); mxv = max(mxv, v0);
#pragma unroll
      for (int k = 1; k < 8; ++k) {
        if (x0 + k >= lim_x) break;
        const float s = grad_sq(o, k);
        smin = fminf(smin, s); smax = fmaxf(smax, s);
        const int v = face ? box_edge(orig, x0 + k, y, z) : o.c[k];
        mnv = min(mnv, v); mxv = max(mxv, v);
      }
      mng = min(mng, f2i(sqrtf(smin))); mxg = max(mxg, f2i(sqrtf(smax)));
    }
  }
  for (int q = 16; q > 0; q >>= 1) {
    mnv = min(mnv, __shfl_xor_sync(0xffffffffu, mnv, q));
    mxv = max(mxv, __shfl_xor_sync(0xffffffffu, mxv, q));
    mng = min(mng, __shfl_xor_sync(0xffffffffu, mng, q));
    mxg = max(mxg, __shfl_xor_sync(0xffffffffu, mxg, q));
  }
  __shared__ int s4[4][VX * VY * VZ / 32];
  const int tid = threadIdx.x + VX * (threadIdx.y + VY * threadIdx.z);
  if ((tid & 31) == 0) { s4[0][tid >> 5] = mnv; s4[1][tid >> 5] = mxv; s4[2][tid >> 5] = mng; s4[3][tid >> 5] = mxg; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < VX * VY * VZ / 32; ++w) {
      mnv = min(mnv, s4[0][w]); mxv = max(mxv, s4[1][w]); mng = min(mng, s4[2][w]); mxg = max(mxg, s4[3][w]);
    }
    atomicMin(stats + 0, mnv); atomicMax(stats + 1, mxv);
    atomicMin(stats + 2, mng); atomicMax(stats + 3, mxg);
  }
}

// fetch_stats in integer arithmetic (speculative; ncu on k_fetch_stats_v8: DRAM traffic exactly 2·N, issue active 80 % — the
// int->float conversions and fp32 products of 24 differences per octet, not the memory system, were the bound).
// (float)dx*(float)dx + ... in fp32 equals the integer dx*dx + dy*dy + dz*dz EXACTLY while every |d| < 4096 and the sum stays
// below 2^24 (all products and partial sums are then representable).  The kernel returns {min v, max v, min S, max S} with the
// integer S; the host accepts them when max(max v, 0) - min(min v, 0) < 4096 (every difference, border zeros included, is
// then below 4096, so S cannot have wrapped either) and max S < 2^24, and takes the square roots itself (IEEE sqrtf on both
// sides); anything else — volumes with neighbouring voxels more than 4095 apart — reruns the fp32 kernel.
// Persistent CTAs of 16 x-octets x 16 rows walk z: a thread keeps the unpacked planes z-1 and z of its octet in registers and
// loads plane z+1 plus the rows y-1 / y+1 of plane z — three 16-byte loads per octet instead of five, and every plane of the
// volume comes from DRAM once (the two earlier versions: one CTA per tile, whose 4 atomics per CTA on the same four words
// serialised in L2 — 0.21 ms whatever the arithmetic cost; then persistent CTAs over scattered tiles, 0.15 ms with the z
// neighbours re-read from DRAM, 390 MB instead of 268).  The x neighbours beyond the octet come from the neighbouring lanes.
#define SX 16
#define SY 16
#define SZC 32  // planes per work item
__global__ void __launch_bounds__(SX* SY) k_fetch_stats_v8i(VolView vol, int32_t* __restrict__ stats, int zlo, int zhi) {
  int mnv = INT32_MAX, mxv = INT32_MIN;
  unsigned mns = 0xFFFFFFFFu, mxs = 0u;
  const unsigned tx = div_up_dev(vol.nx, SX * 8), ty = div_up_dev(vol.ny, SY), tz = div_up_dev(zhi - zlo, SZC);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const size_t sy = vol.nx, sz = (size_t)vol.nx * vol.ny;
  const unsigned lane = (threadIdx.x + SX * threadIdx.y) & 31;  // a warp = 16 octets of row y and 16 of row y+1
  for (unsigned t = blockIdx.x; t < tx * ty * tz; t += gridDim.x) {
    const int x0 = (int)(((t % tx) * SX + threadIdx.x) * 8);
    const int y = (int)(((t / tx) % ty) * SY + threadIdx.y);
    const int z0 = zlo + (int)(t / (tx * ty)) * SZC, z1 = min(z0 + SZC, zhi);
    const bool in = x0 < vol.nx && y < vol.ny;
    const size_t col = (size_t)y * vol.nx + x0;
    int pm[8], pc[8];  // planes z-1 and z of this octet, unpacked
    {
      const uint4 a = (in && z0 > 0) ? __ldg(reinterpret_cast<const uint4*>(vol.v + col + sz * (size_t)(z0 - 1))) : zero;
      const uint4 b = in ? __ldg(reinterpret_cast<const uint4*>(vol.v + col + sz * (size_t)z0)) : zero;
      unpack8(a, pm); unpack8(b, pc);
    }
    for (int z = z0; z < z1; ++z) {
      const size_t row = col + sz * (size_t)z;
      const uint4 qzp = (in && z + 1 < vol.nz) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sz)) : zero;
      const uint4 qym = (in && y > 0) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row - sy)) : zero;
      const uint4 qyp = (in && y + 1 < vol.ny) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sy)) : zero;
      // x neighbours beyond the octet: the neighbouring lanes' edge voxels; the lanes at the ends of the CTA's row segment load theirs
      int xl = __shfl_up_sync(0xffffffffu, pc[7], 1), xr = __shfl_down_sync(0xffffffffu, pc[0], 1);
      if ((lane & 15) == 0) xl = (in && x0 > 0) ? (int)__ldg(vol.v + row - 1) : 0;
      if ((lane & 15) == 15) xr = (in && x0 + 8 < vol.nx) ? (int)__ldg(vol.v + row + 8) : 0;
      if (x0 + 8 >= vol.nx) xr = 0;  // the lane to the right, if any, is outside the volume
      int pp[8];
      unpack8(qzp, pp);
      if (in) {
        const unsigned wym[4] = {qym.x, qym.y, qym.z, qym.w}, wyp[4] = {qyp.x, qyp.y, qyp.z, qyp.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int sh = 16 * (k & 1);
          const int dx = (k == 7 ? xr : pc[k + 1]) - (k == 0 ? xl : pc[k - 1]);
          const int dy = (int)(short)(wyp[k >> 1] >> sh) - (int)(short)(wym[k >> 1] >> sh);
          const int dz = pp[k] - pm[k];
          const unsigned S = (unsigned)(dx * dx) + (unsigned)(dy * dy) + (unsigned)(dz * dz);
          mns = min(mns, S); mxs = max(mxs, S);
          mnv = min(mnv, pc[k]); mxv = max(mxv, pc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { pm[k] = pc[k]; pc[k] = pp[k]; }
    }
  }
  for (int q = 16; q > 0; q >>= 1) {
    mnv = min(mnv, __shfl_xor_sync(0xffffffffu, mnv, q));
    mxv = max(mxv, __shfl_xor_sync(0xffffffffu, mxv, q));
    mns = min(mns, __shfl_xor_sync(0xffffffffu, mns, q));
    mxs = max(mxs, __shfl_xor_sync(0xffffffffu, mxs, q));
  }
  __shared__ unsigned s4[4][SX * SY / 32];
  const int tid = threadIdx.x + SX * threadIdx.y;
  if ((tid & 31) == 0) { s4[0][tid >> 5] = (unsigned)mnv; s4[1][tid >> 5] = (unsigned)mxv; s4[2][tid >> 5] = mns; s4[3][tid >> 5] = mxs; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < SX * SY / 32; ++w) {
      mnv = min(mnv, (int)s4[0][w]); mxv = max(mxv, (int)s4[1][w]); mns = min(mns, s4[2][w]); mxs = max(mxs, s4[3][w]);
    }
    atomicMin(stats + 0, mnv); atomicMax(stats + 1, mxv);
    atomicMin(reinterpret_cast<unsigned*>(stats) + 2, mns); atomicMax(reinterpret_cast<unsigned*>(stats) + 3, mxs);
  }
}

// pin4[8] tells vrk_fetch_stats_finalize what pin4[0..3] will hold: 0 = the final stats, 1 = {min v, max v, min S, max S} of the
// integer kernel (to be checked and converted)
int vrk_fetch_stats_enqueue(vr_ctx* ctx, cudaStream_t stream, const int16_t* vol, int nx, int ny, int nz, int zlo, int zhi,
                            int32_t* dev4, int32_t* pin4) {
  const bool fast = nx % 8 == 0;
  const int32_t init[4] = {INT32_MAX, INT32_MIN, fast ? -1 : INT32_MAX, fast ? 0 : INT32_MIN};  // reference_volume.cpp:22-28 (S: unsigned)
  memcpy(pin4, init, sizeof(init));
  pin4[8] = fast ? 1 : 0;
  VR_CUDA(cudaMemcpyAsync(dev4, pin4, sizeof(init), cudaMemcpyHostToDevice, stream));
  VolView v{vol, nx, ny, nz};
  if (fast) {
    const size_t items = (size_t)div_up(nx, SX * 8) * div_up(ny, SY) * div_up(zhi - zlo, SZC);
    dim3 grid((unsigned)std::min<size_t>(items, (size_t)ctx->sm_count * 8)), block(SX, SY, 1);
    k_fetch_stats_v8i<<<grid, block, 0, stream>>>(v, dev4, zlo, zhi);
  } else {
    dim3 grid(div_up(nx, TX), div_up(ny, TY), div_up(nz, TZ)), block(TX, TY, TZ);
    k_fetch_stats<<<grid, block, 0, stream>>>(v, dev4, zlo, zhi);
  }
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  VR_CUDA(cudaMemcpyAsync(pin4, dev4, sizeof(init), cudaMemcpyDeviceToHost, stream));
  return VR_OK;
}

// after the transfer into pin4 has completed: accept the integer kernel's result or rerun in fp32 (on the context's stream, blocking)
int vrk_fetch_stats_finalize(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int zlo, int zhi, const int32_t* pin4, int32_t out[4]) {
  if (pin4[8] == 0) { memcpy(out, pin4, 4 * sizeof(int32_t)); return VR_OK; }
  const int mnv = pin4[0], mxv = pin4[1];
  const unsigned mns = (unsigned)pin4[2], mxs = (unsigned)pin4[3];
  const bool empty = mnv > mxv;  // no voxel in [zlo, zhi)
  if (empty) { out[0] = out[2] = INT32_MAX; out[1] = out[3] = INT32_MIN; return VR_OK; }
  if ((long long)std::max(mxv, 0) - (long long)std::min(mnv, 0) < 4096 && mxs < (1u << 24)) {
    out[0] = mnv; out[1] = mxv;
    out[2] = (int)sqrtf((float)mns); out[3] = (int)sqrtf((float)mxs);   // implicit float->int of atomic_min/max(int*, float)
    return VR_OK;
  }
  const int32_t init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
  int32_t* pin = ctx->scratch_host + 32;  // bytes 128.. of the pinned scratch
  int32_t* dev = ctx->scratch + 32;
  memcpy(pin, init, sizeof(init));
  VR_CUDA(cudaMemcpyAsync(dev, pin, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  VolView v{vol, nx, ny, nz};
  const size_t tiles = (size_t)div_up(nx, VX * 8) * div_up(ny, VY) * div_up(zhi, VZ);
  dim3 grid((unsigned)std::min<size_t>(tiles, (size_t)ctx->sm_count * 8)), block(VX, VY, VZ);
  k_fetch_stats_v8<false><<<grid, block, 0, ctx->stream>>>(v, v, nx, ny, dev, zlo, zhi);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  VR_CUDA(cudaMemcpyAsync(pin, dev, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, pin, sizeof(init));
  return VR_OK;
}

// under VR_SAMPLING_HW_LINEAR: box = the box-averaged volume (px x (ny+1) x (nz+1)), vol = the volume itself
int vrk_fetch_stats_linear(vr_ctx* ctx, const int16_t* box, int px, const int16_t* vol, int nx, int ny, int nz, int32_t out[4], int zlo,
                           int zhi) {
  const int32_t init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
  memcpy(ctx->scratch_host, init, sizeof(init));
  VR_CUDA(cudaMemcpyAsync(ctx->scratch, ctx->scratch_host, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  VolView b{box, px, ny + 1, nz + 1}, v{vol, nx, ny, nz};
  const int zh = std::min(zhi, nz);
  const size_t tiles = (size_t)div_up(nx, VX * 8) * div_up(ny, VY) * div_up(zh, VZ);
  dim3 grid((unsigned)std::min<size_t>(tiles, (size_t)ctx->sm_count * 8)), block(VX, VY, VZ);
  k_fetch_stats_v8<true><<<grid, block, 0, ctx->stream>>>(b, v, nx, ny, ctx->scratch, zlo, zh);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  VR_CUDA(cudaMemcpyAsync(ctx->scratch_host, ctx->scratch, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->scratch_host, sizeof(init));
  return VR_OK;
}

int vrk_fetch_stats(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int32_t out[4], int zlo, int zhi) {
  VR_TRY(vrk_fetch_stats_enqueue(ctx, ctx->stream, vol, nx, ny, nz, zlo, zhi, ctx->scratch, ctx->scratch_host));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  return vrk_fetch_stats_finalize(ctx, vol, nx, ny, nz, zlo, zhi, ctx->scratch_host, out);
}

// ---- apply_clip: reference_volume_clip.cl:4-15 ---------------------------------------------------------------
// Pure copy (4 bytes per output voxel).  A thread produces 8 consecutive output voxels and writes them with one 16-byte
// store (the destination is a fresh allocation, so 16-byte aligned; output rows are not, so a group may straddle rows: the
// (x,y,z) of its first voxel comes from one 32-bit division pair, the rest by carry).  Source rows start at arbitrary
// offsets, so the reads are 2-byte loads — consecutive lanes read consecutive addresses, L1 merges them into full sectors.
// Reads outside the source are border reads (0), as the reference's samplerless read_imagei on a CLK_ADDRESS_CLAMP-less
// image is only ever used in range (reference_volume.cpp:57-59 asserts it) — kept defined here.
__global__ void __launch_bounds__(256) k_clip(VolView src, int sx, int sy, int sz, int16_t* __restrict__ dst, int nx,
                                              int ny, unsigned n) {
  const unsigned ngroups = (n + 7u) >> 3;
  for (unsigned gi = blockIdx.x * blockDim.x + threadIdx.x; gi < ngroups; gi += gridDim.x * blockDim.x) {
    const unsigned i0 = gi << 3;
    int x = (int)(i0 % (unsigned)nx);
    const unsigned t = i0 / (unsigned)nx;
    int y = (int)(t % (unsigned)ny), z = (int)(t / (unsigned)ny);
    unsigned v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = (unsigned)(unsigned short)src.at(sx + x, sy + y, sz + z);  // the last group may run past n: those reads are in range or 0
      if (++x == nx) { x = 0; if (++y == ny) { y = 0; ++z; } }
    }
    if (i0 + 8u <= n) {
      *reinterpret_cast<uint4*>(dst + i0) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
    } else {
      for (unsigned k = 0; i0 + k < n; ++k) dst[i0 + k] = (int16_t)v[k];
    }
  }
}

int vrk_clip(vr_ctx* ctx, const int16_t* src, int snx, int sny, int snz, const uint32_t start[3], int16_t* dst, int nx,
             int ny, int nz) {
  VolView v{src, snx, sny, snz};
  const size_t n = (size_t)nx * ny * nz;  // < 2^32 - 1: the output is a sub-box of an uploaded volume or at most as large
  VR_REQUIRE(n < ((size_t)1 << 32) - 8, "vr_volume_clip: more than 2^32-9 voxels");
  unsigned blocks = (unsigned)std::min<size_t>(div_up(div_up(n, 8), 256), (size_t)ctx->sm_count * 16);
  k_clip<<<blocks, 256, 0, ctx->stream>>>(v, (int)start[0], (int)start[1], (int)start[2], dst, nx, ny, (unsigned)n);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- bilateral_filter: volume_filter.cl:5-11 + utility_filter.cl:38-62 ---------------------------------------
// 125 taps per voxel, one exp per tap in the reference.  The block stages its (TX+4)x(TY+4)x(TZ+4) neighbourhood in shared
// memory once so every tap is an LDS.  pow(d, 2.0f) is evaluated as d*d.
// The weight exp(-posd - cold) takes few distinct values: posd depends on dx^2+dy^2+dz^2 (10 classes for a radius of 2), and the
// voxels are integers, so cold = a*a/2 with a = |mid - local|; for a >= 16 the argument is <= -128 and expf underflows to
// exactly 0.  Every block evaluates the 10 x 17 weights once with the SAME fp32 expression and the taps look them up:
// bit-identical output, 125 LDS instead of 125 expf per voxel (9.2 -> see profiles/ at 512^3).
#define BIL_A 17
__device__ __forceinline__ int bil_class(int d2) {  // 0,1,2,3,4,5,6,8,9,12 -> 0..9
  return d2 <= 6 ? d2 : (d2 == 8 ? 7 : (d2 == 9 ? 8 : 9));
}
// A thread computes TWO outputs, z and z + 1: the tile value of a tap is loaded once for both (6 planes instead of 2 x 5), and
// the index of the weight comes without a float -> int conversion (ncu on the first version: XU pipe 63 % busy with 125 F2I per
// voxel, issue active 82 %): |mid - local| is an integer-valued float, fma(min(|d|, 16), 4, 1.5 * 2^23) has the byte offset 4 a in
// its low mantissa bits.  Each output still accumulates its 125 taps in the reference's order (dz, dy, dx ascending).
// dnx/dny/dnz: dims of the output (== the volume's; under VR_SAMPLING_HW_LINEAR `vol` is the box-averaged volume, one larger)
__global__ void __launch_bounds__(TX* TY* TZ) k_bilateral(VolView vol, int16_t* __restrict__ dst, int dnx, int dny, int dnz) {
  __shared__ float tile[2 * TZ + 4][TY + 4][TX + 4];
  __shared__ float wtab[10 * BIL_A];
  const int bx = blockIdx.x * TX, by = blockIdx.y * TY, bz = blockIdx.z * (2 * TZ);
  const int tid = threadIdx.x + TX * (threadIdx.y + TY * threadIdx.z);
  const float sigmas = 0.6f, sigmar = 1.0f;
  if (tid < 10 * BIL_A) {
    const int cls = tid / BIL_A, a = tid % BIL_A;
    const int d2 = cls <= 6 ? cls : (cls == 7 ? 8 : (cls == 8 ? 9 : 12));
    const float posd = ((float)d2) / (2 * sigmas * sigmas);
    const float diff = (float)a;
    const float cold = (diff * diff) / (2 * sigmar * sigmar);
    wtab[tid] = a == BIL_A - 1 ? 0.0f : expf(-posd - cold);  // a >= 16: expf(<= -128) == 0
  }
  for (int i = tid; i < (2 * TZ + 4) * (TY + 4) * (TX + 4); i += TX * TY * TZ) {
    int lx = i % (TX + 4);
    int t = i / (TX + 4);
    int ly = t % (TY + 4);
    int lz = t / (TY + 4);
    tile[lz][ly][lx] = (float)vol.at(bx + lx - 2, by + ly - 2, bz + lz - 2);
  }
  __syncthreads();
  const int x = bx + threadIdx.x, y = by + threadIdx.y, z = bz + 2 * threadIdx.z;
  if (x >= dnx || y >= dny || z >= dnz) return;
  const int lz0 = 2 * threadIdx.z + 2;  // tile plane of output z
  const float mid0 = tile[lz0][threadIdx.y + 2][threadIdx.x + 2], mid1 = tile[lz0 + 1][threadIdx.y + 2][threadIdx.x + 2];
  float out0 = 0.0f, wp0 = 0.0f, out1 = 0.0f, wp1 = 0.0f;
  const char* wbytes = reinterpret_cast<const char*>(wtab);
  auto weight = [&](float mid, float local, int cls) {
    const float t = __fmaf_rn(fminf(fabsf(mid - local), (float)(BIL_A - 1)), 4.0f, 12582912.0f);  // exact: both are integers of |value| < 2^15
    return *reinterpret_cast<const float*>(wbytes + cls * (BIL_A * 4) + (__float_as_int(t) - 0x4B400000));
  };
#pragma unroll
  for (int pz = -2; pz <= 3; ++pz)  // tile plane lz0 + pz: tap dz = pz of output z, tap dz = pz - 1 of output z + 1
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
      for (int dx = -2; dx <= 2; ++dx) {
        const float local = tile[lz0 + pz][threadIdx.y + 2 + dy][threadIdx.x + 2 + dx];
        if (pz <= 2) {
          const float w = weight(mid0, local, bil_class(dx * dx + dy * dy + pz * pz));
          wp0 += w;
          out0 += local * w;
        }
        if (pz >= -1) {
          const float w = weight(mid1, local, bil_class(dx * dx + dy * dy + (pz - 1) * (pz - 1)));
          wp1 += w;
          out1 += local * w;
        }
      }
  const size_t o = (size_t)x + (size_t)dnx * ((size_t)y + (size_t)dny * (size_t)z);
  dst[o] = (int16_t)f2s(out0 / wp0);
  if (z + 1 < dnz) dst[o + (size_t)dnx * dny] = (int16_t)f2s(out1 / wp1);
}

int vrk_bilateral(vr_ctx* ctx, const int16_t* src, int16_t* dst, int nx, int ny, int nz) {
  VolView v{src, nx, ny, nz};
  dim3 grid(div_up(nx, TX), div_up(ny, TY), div_up(nz, 2 * TZ)), block(TX, TY, TZ);
  k_bilateral<<<grid, block, 0, ctx->stream>>>(v, dst, nx, ny, nz);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
// centre and taps from the box-averaged volume (border addressing: utility_filter.cl:40), output for the volume's own voxels
int vrk_bilateral_linear(vr_ctx* ctx, const int16_t* box, int px, int nx, int ny, int nz, int16_t* dst) {
  VolView b{box, px, ny + 1, nz + 1};
  dim3 grid(div_up(nx, TX), div_up(ny, TY), div_up(nz, 2 * TZ)), block(TX, TY, TZ);
  k_bilateral<<<grid, block, 0, ctx->stream>>>(b, dst, nx, ny, nz);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- tf_sort_values: histogram.cl:4-32 ------------------------------------------------------------------------
// Flat index x*height + y computed exactly as written; indices outside [0, W*H) are dropped (the reference
// writes out of bounds there); the y==H aliasing into the next column is kept.
// Lanes of a warp that land in the same bin are merged with __match_any_sync so a hot bin costs one atomic
// per warp instead of 32.
__global__ void __launch_bounds__(TX* TY* TZ) k_histogram(VolView vol, uint32_t* __restrict__ bins, int width,
                                                          int height, float min_v, float max_v, float min_g,
                                                          float max_g, int zlo, int zhi) {
  const int x = blockIdx.x * TX + threadIdx.x;
  const int y = blockIdx.y * TY + threadIdx.y;
  const int z = blockIdx.z * TZ + threadIdx.z;
  long long flat = -1;
  if (x < vol.nx && y < vol.ny && z >= zlo && z < zhi) {
    int ref_value = vol.at(x, y, z);
    float g = length3(gradient_voxel(vol, x, y, z));
    if (!(g > max_g) && !((float)ref_value > max_v)) {
      float value_range = max_v - min_v;
      float gradient_range = max_g - min_g;
      int px = f2i(roundf((((float)ref_value - min_v) / value_range) * (float)width));
      int py = f2i(roundf(((g - min_g) / gradient_range) * (float)height));
      flat = (long long)px * height + py;
      if (flat < 0 || flat >= (long long)width * height) flat = -1;
    }
  }
  const unsigned active = __ballot_sync(0xffffffffu, flat >= 0);
  if (flat >= 0) {
    const unsigned peers = __match_any_sync(active, (int)flat);
    const int leader = __ffs(peers) - 1;
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(bins + flat, (uint32_t)__popc(peers));
  }
}

// tf_sort_values, vectorised front end (nx % 8 == 0): same binning arithmetic per voxel, 8 voxels per thread.
// CT-like volumes put most voxels into a few hundred neighbouring bins (air / soft tissue, small gradients), and same-address
// atomics serialise in L2 (the first version spent 3.2 ms at 512^3 on them; a shared-memory hash table with __match_any
// merging 2.2 ms).  Persistent CTAs now count into a direct-mapped WINDOW of HWX x HWY bins in shared memory (plain shared
// atomics, no matching, no probing) placed at the bin of (smallest value, gradient 0); voxels that fall outside the window
// go to global memory directly; every CTA flushes its window once at the end.  Bit-identical counts: integer adds commute.
#define HWX 64
#define HWY 128
template <bool LINEAR>
__global__ void __launch_bounds__(VX* VY* VZ) k_histogram_v8(VolView vol, VolView orig, int lim_x, int lim_y, uint32_t* __restrict__ bins,
                                                             int width, int height, float min_v, float max_v, float min_g, float max_g,
                                                             int zlo, int zhi, int wx0, int wy0) {
  __shared__ unsigned win[HWX * HWY];
  const int tid = threadIdx.x + VX * (threadIdx.y + VY * threadIdx.z);
  for (int i = tid; i < HWX * HWY; i += VX * VY * VZ) win[i] = 0;
  __syncthreads();
  const unsigned tx = div_up_dev(lim_x, VX * 8), ty = div_up_dev(lim_y, VY), tz = div_up_dev(zhi, VZ);
  const float value_range = max_v - min_v, gradient_range = max_g - min_g;
  const long long nbins = (long long)width * height;
  for (unsigned t = blockIdx.x; t < tx * ty * tz; t += gridDim.x) {
    const int x0 = (int)(((t % tx) * VX + threadIdx.x) * 8);
    const int y = (int)(((t / tx) % ty) * VY + threadIdx.y);
    const int z = (int)((t / (tx * ty)) * VZ + threadIdx.z);
    if (!(x0 < lim_x && y < lim_y && z >= zlo && z < zhi)) continue;
    const Octet o = load_octet(vol, x0, y, z);
    const bool face = LINEAR && (y == 0 || z == 0);
    // a thread's 8 voxels often share a bin (air, the inside of a homogeneous object): run-length encode them, and merge
    // equal out-of-window bins across the warp, so that a hot bin anywhere in the grid costs one global atomic per warp
    long long run = -1;
    unsigned run_n = 0;
    int run_w = -1;
    auto flush = [&]() {
      if (!run_n) return;
      if (run_w >= 0) atomicAdd(&win[run_w], run_n);
      else {
        const unsigned peers = __match_any_sync(__activemask(), run);
        const unsigned total = __reduce_add_sync(peers, run_n);
        if ((tid & 31) == __ffs(peers) - 1) atomicAdd(bins + run, total);
      }
    };
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (x0 + k >= lim_x) break;
      const float g = sqrtf(grad_sq(o, k));
      const int value = (LINEAR && (face || x0 + k == 0)) ? box_edge(orig, x0 + k, y, z) : o.c[k];
      long long flat = -1;
      int w = -1;
      if (!(g > max_g || (float)value > max_v)) {
        const int px = f2i(roundf((((float)value - min_v) / value_range) * (float)width));
        const int py = f2i(roundf(((g - min_g) / gradient_range) * (float)height));
        flat = (long long)px * height + py;
        if (flat < 0 || flat >= nbins) flat = -1;
        const unsigned wx = (unsigned)(px - wx0), wy = (unsigned)(py - wy0);
        if (flat >= 0 && wx < HWX && wy < HWY && py < height) w = (int)(wx * HWY + wy);
      }
      if (flat != run) {
        flush();
        run = flat; run_w = w; run_n = 0;
      }
      if (flat >= 0) ++run_n;
    }
    flush();
  }
  __syncthreads();
  for (int i = tid; i < HWX * HWY; i += VX * VY * VZ) {
    const unsigned c = win[i];
    if (c) atomicAdd(bins + (long long)(wx0 + i / HWY) * height + (wy0 + i % HWY), c);  // non-zero only for bins inside the grid
  }
}

// tf_sort_values through tables.  ncu on k_histogram_v8 (r2c): 71 % issue active, ~90 instructions per voxel — IEEE sqrt, two IEEE
// divisions and two roundf per voxel, which bit-exactness fixes.  But a voxel's column px is a function of its VALUE alone and its
// row py of its squared gradient S alone, and both arguments are small integers for real data: when every difference of
// neighbouring voxels (border zeros included) is below 4096 and max_g < 4096, S = dx^2 + dy^2 + dz^2 is the same number in int32
// and in the fp32 evaluation order of length() (every product and partial sum below 2^24 is exact; a sum that is not exact is
// above max_g^2 in both), so k_hist_tables evaluates the reference's expressions ONCE per distinct value (<= 4096) and per
// S < HPY, and the volume pass looks them up in shared memory: ~25 instructions per voxel.  S >= HPY (steep edges) takes the
// expressions directly.  Volumes outside the guard use k_histogram_v8.  The pass itself is k_fetch_stats_v8i's: persistent CTAs
// walk z with the unpacked planes z-1 and z in registers, three 16-byte loads per octet.
#define HPX 4096   // value table entries
#define HP1 8192   // squared-gradient table, one entry per S < HP1 (gradient < 90.5)
#define HP2 8192   // one entry per block of 128 S values up to S < 2^20 (gradient < 1024): the row if the whole block shares it
#define HSKIP (-32768)   // the voxel is not counted (value > max_v / gradient > max_g)
#define HNONUNI (-32767) // the block of S values spans two rows (or the skip threshold): evaluate the expression
struct HistArgs {
  int width, height;
  float min_v, max_v, min_g, max_g;
  int vmin, nval;  // table covers values vmin .. vmin + nval - 1
};
// the reference's expressions (histogram.cl:17-27); the host has checked that every result fits 16 bits with room for the markers
__device__ __forceinline__ int hist_px(const HistArgs& a, int value) {
  if ((float)value > a.max_v) return HSKIP;
  return f2i(roundf((((float)value - a.min_v) / (a.max_v - a.min_v)) * (float)a.width));
}
__device__ __forceinline__ int hist_py(const HistArgs& a, unsigned S) {
  const float g = sqrtf((float)S);
  if (g > a.max_g) return HSKIP;
  return f2i(roundf(((g - a.min_g) / (a.max_g - a.min_g)) * (float)a.height));
}
__global__ void __launch_bounds__(256) k_hist_tables(HistArgs a, short* __restrict__ pxl, short* __restrict__ pyl1, short* __restrict__ pyl2) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HPX + HP1 + HP2; i += gridDim.x * blockDim.x) {
    if (i < HPX) pxl[i] = (short)(i < a.nval ? hist_px(a, a.vmin + i) : HSKIP);
    else if (i < HPX + HP1) pyl1[i - HPX] = (short)hist_py(a, (unsigned)(i - HPX));
    else {
      const unsigned b = (unsigned)(i - HPX - HP1);
      const int p0 = hist_py(a, b << 7), p1 = hist_py(a, (b << 7) + 127u);  // monotone in S: equal ends = one row for the block
      pyl2[b] = (short)(p0 == p1 ? p0 : HNONUNI);
    }
  }
}
// One CTA of 16 x HBY threads per SM with (almost) all of its shared memory: the tables and a window of wc columns x wr rows of
// the bin grid (HWBINS bins) placed at (column of the smallest value, row of gradient 0) and shaped by the host to span the
// data's columns — 500 x 90 bins for the bench volume's 500 x 500 grid, 93 % of its voxels (the 64 x 128 window of
// k_histogram_v8 caught a fraction of them; the rest went to global atomics one warp-merged bin at a time, and THAT was its
// time, not the arithmetic: the table look-ups alone changed nothing, 1.06 ms both).
#define HBY 64
#define HZC 16       // planes per work item
#define HWBINS 45056 // window bins (176 KiB)
// LINEAR: `vol` is the box-averaged volume (its apron rows / columns / planes are gradient neighbours only), `orig` the volume
// itself: voxels x < lim_x, y < lim_y count, and the value read on the three low faces is box_edge's (see k_boxavg).
template <bool LINEAR>
__global__ void __launch_bounds__(SX* HBY, 1) k_histogram_lut(VolView vol, VolView orig, int lim_x, int lim_y, uint32_t* __restrict__ bins,
                                                              HistArgs a, int zlo, int zhi, int wx0, int wy0, int wc, int wr,
                                                              const short* __restrict__ g_tables) {
  extern __shared__ unsigned hsm[];
  unsigned* win = hsm;                                     // wc x wr window of the bin grid
  short* pxl = reinterpret_cast<short*>(hsm + HWBINS);     // HPX | HP1 | HP2, as k_hist_tables wrote them
  const short* pyl1 = pxl + HPX;
  const short* pyl2 = pyl1 + HP1;
  const int tid = threadIdx.x + SX * threadIdx.y;
  const int wbins = wc * wr;
  for (int i = tid; i < wbins; i += SX * HBY) win[i] = 0;
  for (int i = tid; i < (HPX + HP1 + HP2) / 2; i += SX * HBY) reinterpret_cast<unsigned*>(pxl)[i] = reinterpret_cast<const unsigned*>(g_tables)[i];
  __syncthreads();
  const unsigned tx = div_up_dev(lim_x, SX * 8), ty = div_up_dev(lim_y, HBY), tz = div_up_dev(zhi - zlo, HZC);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  const size_t sy = vol.nx, sz = (size_t)vol.nx * vol.ny;
  const unsigned lane = tid & 31;  // a warp = 16 octets of row y and 16 of row y+1
  const long long nbins = (long long)a.width * a.height;
  for (unsigned t = blockIdx.x; t < tx * ty * tz; t += gridDim.x) {
    const int x0 = (int)(((t % tx) * SX + threadIdx.x) * 8);
    const int y = (int)(((t / tx) % ty) * HBY + threadIdx.y);
    const int z0 = zlo + (int)(t / (tx * ty)) * HZC, z1 = min(z0 + HZC, zhi);
    const bool in = x0 < vol.nx && y < vol.ny;     // the octet exists (is loaded: its edge voxels are other lanes' neighbours)
    const bool cnt = x0 < lim_x && y < lim_y;      // ... and (some of) its voxels are counted
    const size_t col = (size_t)y * vol.nx + x0;
    int pm[8], pc[8];  // planes z-1 and z of this octet, unpacked
    {
      const uint4 qa = (in && z0 > 0) ? __ldg(reinterpret_cast<const uint4*>(vol.v + col + sz * (size_t)(z0 - 1))) : zero;
      const uint4 qb = in ? __ldg(reinterpret_cast<const uint4*>(vol.v + col + sz * (size_t)z0)) : zero;
      unpack8(qa, pm); unpack8(qb, pc);
    }
    for (int z = z0; z < z1; ++z) {
      const size_t row = col + sz * (size_t)z;
      const uint4 qzp = (in && z + 1 < vol.nz) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sz)) : zero;
      const uint4 qym = (in && y > 0) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row - sy)) : zero;
      const uint4 qyp = (in && y + 1 < vol.ny) ? __ldg(reinterpret_cast<const uint4*>(vol.v + row + sy)) : zero;
      int xl = __shfl_up_sync(0xffffffffu, pc[7], 1), xr = __shfl_down_sync(0xffffffffu, pc[0], 1);
      if ((lane & 15) == 0) xl = (in && x0 > 0) ? (int)__ldg(vol.v + row - 1) : 0;
      if ((lane & 15) == 15) xr = (in && x0 + 8 < vol.nx) ? (int)__ldg(vol.v + row + 8) : 0;
      if (x0 + 8 >= vol.nx) xr = 0;  // the lane to the right, if any, is outside the volume
      int pp[8];
      unpack8(qzp, pp);
      // one shared-memory add per voxel inside the window (one per warp where all 32 lanes hit the same bin: air, the inside of
      // a homogeneous object), one RED.ADD to the grid outside it — no warp-wide matching: steep-edge voxels are few per bin but
      // present in almost every warp, and a match per voxel is what k_histogram_v8 spent its time on
      const unsigned wym[4] = {qym.x, qym.y, qym.z, qym.w}, wyp[4] = {qyp.x, qyp.y, qyp.z, qyp.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int sh = 16 * (k & 1);
        const int dx = (k == 7 ? xr : pc[k + 1]) - (k == 0 ? xl : pc[k - 1]);
        const int dy = (int)(short)(wyp[k >> 1] >> sh) - (int)(short)(wym[k >> 1] >> sh);
        const int dz = pp[k] - pm[k];
        const unsigned S = (unsigned)(dx * dx) + (unsigned)(dy * dy) + (unsigned)(dz * dz);
        int value = pc[k];
        if (LINEAR && (y == 0 || z == 0 || x0 + k == 0) && cnt && x0 + k < lim_x) value = box_edge(orig, x0 + k, y, z);
        const unsigned vi = (unsigned)(value - a.vmin);
        const int px = vi < (unsigned)a.nval ? (int)pxl[vi] : hist_px(a, value);
        int py;
        if (S < HP1) py = pyl1[S];
        else {
          py = (S >> 7) < HP2 ? (int)pyl2[S >> 7] : HNONUNI;
          if (py == HNONUNI) py = hist_py(a, S);
        }
        // inside the window (columns wx0 .. wx0 + wc - 1 lie inside the grid, so with py < height the bin does too; the markers
        // are large negative numbers and fail the unsigned tests): no 64-bit bin number needed
        const unsigned wx = (unsigned)(px - wx0), wy = (unsigned)(py - wy0);
        const bool counted = cnt && (!LINEAR || x0 + k < lim_x);
        const int w = (counted && wx < (unsigned)wc && wy < (unsigned)wr && py < a.height) ? (int)(wx * (unsigned)wr + wy) : -1;
        int same;
        __match_all_sync(0xffffffffu, w, &same);
        if (same && w >= 0) {
          if (lane == 0) atomicAdd(&win[w], 32u);
        } else if (w >= 0) {
          atomicAdd(&win[w], 1u);
        } else if (counted && px != HSKIP && py != HSKIP) {
          const long long flat = (long long)px * a.height + py;
          if (flat >= 0 && flat < nbins) atomicAdd(bins + flat, 1u);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) { pm[k] = pc[k]; pc[k] = pp[k]; }
    }
  }
  __syncthreads();
  for (int i = tid; i < wbins; i += SX * HBY) {
    const unsigned c = win[i];
    if (c) atomicAdd(bins + (long long)(wx0 + i / wr) * a.height + (wy0 + i % wr), c);  // non-zero only for bins inside the grid
  }
}

// the table path of tf_sort_values, if its guards hold: every difference of two voxels (or a voxel and a border zero) below 4096,
// no counted gradient above 4095, and every column / row number the expressions can produce within 16 bits.  [vmin, vmax] must
// bound every value `vol` holds.  Returns false when the caller has to use k_histogram_v8.
template <bool LINEAR>
static bool hist_tables_launch(vr_ctx* ctx, const VolView& vol, const VolView& orig, int nx, int ny, int nz, int width, int height,
                               const float range[4], uint32_t* bins_dev, int zlo, int zhi, int vmin, int vmax, int wx0, int wy0) {
  const float vr = range[1] - range[0], gr = range[3] - range[2];
  const long long span = (long long)std::max(vmax, 0) - (long long)std::min(vmin, 0);
  if (!(vmax >= vmin && span < 4096 && range[3] < 4096.0f && zhi > zlo && vr > 0.0f && gr > 0.0f)) return false;
  const float px_lo = ((float)vmin - range[0]) / vr * (float)width, px_hi = ((float)vmax - range[0]) / vr * (float)width;
  const float py_lo = (0.0f - range[2]) / gr * (float)height, py_hi = (range[3] - range[2]) / gr * (float)height;
  if (!(fabsf(px_lo) < 32000.0f && fabsf(px_hi) < 32000.0f && fabsf(py_lo) < 32000.0f && fabsf(py_hi) < 32000.0f)) return false;
  HistArgs a{width, height, range[0], range[1], range[2], range[3], vmin, std::min(vmax - vmin + 1, HPX)};
  short* tables = nullptr;
  if (cudaMallocAsync(reinterpret_cast<void**>(&tables), (HPX + HP1 + HP2) * sizeof(short), ctx->stream) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  k_hist_tables<<<(HPX + HP1 + HP2) / 256, 256, 0, ctx->stream>>>(a, tables, tables + HPX, tables + HPX + HP1);
  const size_t smem = (size_t)HWBINS * 4 + (size_t)(HPX + HP1 + HP2) * 2;
  // per device (a host may drive several GPUs from one process): set on every call, it is cheap
  cudaFuncSetAttribute(k_histogram_lut<LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // window: the columns the data's values span (as many as leave 16 rows), then as many rows as fit
  const float top_v = std::min((float)vmax, range[1]);
  const int c_hi = std::max(0, std::min(width - 1, (int)roundf((top_v - range[0]) / vr * (float)width)));
  const int ncols = std::max(1, c_hi - wx0 + 1);
  const int wr = std::max(16, std::min(height - wy0, HWBINS / ncols));
  const int wc = std::max(1, std::min(ncols, HWBINS / wr));
  const int zh = std::min(zhi, nz);
  const size_t items = (size_t)div_up(nx, SX * 8) * div_up(ny, HBY) * div_up(zh - zlo, HZC);
  k_histogram_lut<LINEAR><<<(unsigned)std::min<size_t>(items, (size_t)ctx->sm_count), dim3(SX, HBY, 1), smem, ctx->stream>>>(
      vol, orig, nx, ny, bins_dev, a, zlo, zh, wx0, wy0, wc, wr, tables);
  ctx->launches += 2;
  cudaFreeAsync(tables, ctx->stream);
  return true;
}

int vrk_histogram(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, int width, int height, const float range[4],
                  uint32_t* bins_dev, int zlo, int zhi, int vol_min_value, int vol_max_value) {
  VR_CUDA(cudaMemsetAsync(bins_dev, 0, sizeof(uint32_t) * (size_t)width * height, ctx->stream));
  VolView v{vol, nx, ny, nz};
  if (nx % 8 == 0) {
    // window origin: the bin of the smallest value at gradient 0 (same arithmetic as the kernel), clamped into the grid
    const float vr = range[1] - range[0], gr = range[3] - range[2];
    int wx0 = 0, wy0 = 0;
    if (vr > 0.0f) wx0 = std::max(0, std::min(width - 1, (int)roundf(((float)vol_min_value - range[0]) / vr * (float)width)));
    if (gr > 0.0f) wy0 = std::max(0, std::min(height - 1, (int)roundf((0.0f - range[2]) / gr * (float)height)));
    if (hist_tables_launch<false>(ctx, v, v, nx, ny, nz, width, height, range, bins_dev, zlo, zhi, vol_min_value, vol_max_value, wx0, wy0))
    {
      VR_CUDA(cudaGetLastError());
      return VR_OK;
    }
    const size_t tiles = (size_t)div_up(nx, VX * 8) * div_up(ny, VY) * div_up(nz, VZ);
    dim3 grid((unsigned)std::min<size_t>(tiles, (size_t)ctx->sm_count * 6)), block(VX, VY, VZ);
    k_histogram_v8<false><<<grid, block, 0, ctx->stream>>>(v, v, nx, ny, bins_dev, width, height, range[0], range[1], range[2], range[3], zlo,
                                                          std::min(zhi, nz), wx0, wy0);
  } else {
    dim3 grid(div_up(nx, TX), div_up(ny, TY), div_up(nz, TZ)), block(TX, TY, TZ);
    k_histogram<<<grid, block, 0, ctx->stream>>>(v, bins_dev, width, height, range[0], range[1], range[2], range[3], zlo, zhi);
  }
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// raw_min / raw_max: bounds of the values of the volume the box average was built from (every box value, apron included, is a
// rounded mean of such values and border zeros, hence inside [min(raw_min, 0), max(raw_max, 0)]); raw_max < raw_min: unknown
int vrk_histogram_linear(vr_ctx* ctx, const int16_t* box, int px, const int16_t* vol, int nx, int ny, int nz, int width, int height,
                         const float range[4], uint32_t* bins_dev, int zlo, int zhi, int vol_min_value, int raw_min, int raw_max) {
  VR_CUDA(cudaMemsetAsync(bins_dev, 0, sizeof(uint32_t) * (size_t)width * height, ctx->stream));
  VolView b{box, px, ny + 1, nz + 1}, v{vol, nx, ny, nz};
  const size_t tiles = (size_t)div_up(nx, VX * 8) * div_up(ny, VY) * div_up(nz, VZ);
  dim3 grid((unsigned)std::min<size_t>(tiles, (size_t)ctx->sm_count * 6)), block(VX, VY, VZ);
  const float vr = range[1] - range[0], gr = range[3] - range[2];
  int wx0 = 0, wy0 = 0;
  if (vr > 0.0f) wx0 = std::max(0, std::min(width - 1, (int)roundf(((float)vol_min_value - range[0]) / vr * (float)width)));
  if (gr > 0.0f) wy0 = std::max(0, std::min(height - 1, (int)roundf((0.0f - range[2]) / gr * (float)height)));
  if (raw_max >= raw_min &&
      hist_tables_launch<true>(ctx, b, v, nx, ny, nz, width, height, range, bins_dev, zlo, zhi, std::min(raw_min, 0), std::max(raw_max, 0), wx0, wy0))
  {
    VR_CUDA(cudaGetLastError());
    return VR_OK;
  }
  k_histogram_v8<true><<<grid, block, 0, ctx->stream>>>(b, v, nx, ny, bins_dev, width, height, range[0], range[1], range[2], range[3], zlo,
                                                        std::min(zhi, nz), wx0, wy0);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- tf_flush_color_frame: histogram.cl:34-69 — see k_tf_color_frame_ranked below ----------------------------------------------
// ---- render_tf on the device -------------------------------------------------------------------------
// renderer.cpp:65-96 pulls the bins to the host, rounds every non-zero count down to two significant digits, collects the
// distinct values in a std::set, pushes bins + set back and lets the colour kernel search the set.  A rounded count is
// (leading one or two digits) x 10^d, so it has a CODE d*100 + digits < 1000 that orders like the value: mark the codes that
// occur, prefix-sum the marks into ranks, colour by rank.  No host round trip, no search.  The integer rounding equals the
// reference's double pow/log10/floor expression for every count < 2^31 (tests/test_oracle_cpu.py checks it).
__device__ __forceinline__ int tf_round_code(int v, int* corrected) {
  int p = 1, d = 0;
  while (v / p >= 100) { p *= 10; ++d; }   // p = max(10^(floor(log10 v) - 1), 1)
  const int lead = v / p;
  *corrected = lead * p;
  return d * 100 + lead;
}
__global__ void __launch_bounds__(256) k_tf_round_mark(int32_t* __restrict__ bins, size_t n, int* __restrict__ flags) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int v = bins[i];
    if (v != 0) {
      int c;
      flags[tf_round_code(v, &c)] = 1;
      bins[i] = c;
    }
  }
}
// one block of 1024 threads: rank[code] = number of marked codes below it; rank[1024] = number of marked codes
__global__ void __launch_bounds__(1024) k_tf_rank(const int* __restrict__ flags, int* __restrict__ rank) {
  __shared__ int s[1024];
  const int t = threadIdx.x;
  const int f = flags[t];
  s[t] = f;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = t >= o ? s[t - o] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  rank[t] = s[t] - f;
  if (t == 1023) rank[1024] = s[t];
}
__global__ void __launch_bounds__(256) k_tf_color_frame_ranked(const int32_t* __restrict__ bins, const int* __restrict__ rank,
                                                               int width, int height, uchar4* __restrict__ out) {
  const int px = blockIdx.x * 16 + (threadIdx.x & 15);
  const int py = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (px >= width || py >= height) return;
  const int len = rank[1024];
  if (len == 0) {  // renderer.cpp:84-86: the colour kernel is skipped, the frame keeps its zero initialisation
    out[(size_t)py * width + px] = make_uchar4(0, 0, 0, 0);
    return;
  }
  const int value = bins[(size_t)px * height + (height - py - 1)];
  int result = 0;
  if (value != 0) {
    int c;
    const int local_value = rank[tf_round_code(value, &c)];
    result = f2i(20.0f + (((float)local_value) / (float)len) * (255.0f - 20.0f));
  }
  const unsigned char g = (unsigned char)max(0, min(255, result));
  out[(size_t)py * width + px] = make_uchar4(g, g, g, 255);
}

// bins (counts) -> RGBA image, all on the device; scratch = 2049 ints
int vrk_tf_image(vr_ctx* ctx, int32_t* bins_dev, int* scratch_dev, int width, int height, uchar4* out_dev) {
  const size_t n = (size_t)width * height;
  VR_CUDA(cudaMemsetAsync(scratch_dev, 0, 1024 * sizeof(int), ctx->stream));
  k_tf_round_mark<<<(unsigned)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(bins_dev, n, scratch_dev);
  k_tf_rank<<<1, 1024, 0, ctx->stream>>>(scratch_dev, scratch_dev + 1024);
  dim3 grid(div_up(width, 16), div_up(height, 16));
  k_tf_color_frame_ranked<<<grid, 256, 0, ctx->stream>>>(bins_dev, scratch_dev + 1024, width, height, out_dev);
  ctx->launches += 3;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- buffer_reset: buffer_reset.cl:3-13 — 8 bytes per voxel, pure store bandwidth ------------------------------
__global__ void __launch_bounds__(256) k_cache_reset(uint4* __restrict__ p, size_t n16, uint32_t* __restrict__ tail,
                                                     int ntail) {
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0;
}

// frame reset when only the primary-hit voxels of the current hit buffer can be non-zero (the path tracer writes the cache at
// primary hits only, ray_marching.cl:39,76): W*H scattered 8-byte stores instead of 8 bytes x voxels
__global__ void __launch_bounds__(256) k_cache_reset_hits(uint2* __restrict__ cache, const uint32_t* __restrict__ hit, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = hit[i];
  if (v != 0xFFFFFFFFu) cache[v] = make_uint2(0u, 0u);
}
int vrk_cache_reset_hits(vr_ctx* ctx, uint32_t* cache, const uint32_t* hit, size_t pixels) {
  k_cache_reset_hits<<<div_up(pixels, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint2*>(cache), hit, pixels);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

int vrk_cache_reset(vr_ctx* ctx, uint32_t* cache, size_t voxels, cudaStream_t stream) {
  if (!stream) stream = ctx->stream;
  const size_t words = voxels * 2;
  const size_t n16 = words / 4;
  const int ntail = (int)(words - n16 * 4);
  unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(div_up(n16, 256), 1), (size_t)ctx->sm_count * 8);
  k_cache_reset<<<blocks, 256, 0, stream>>>(reinterpret_cast<uint4*>(cache), n16, cache + n16 * 4, ntail);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- position-weighted 64-bit checksum of a device buffer (instrumentation: multi-GPU results are compared at full size without
// moving gigabytes to the host) ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_checksum(const uint32_t* __restrict__ words, size_t nwords, const uint8_t* __restrict__ tail,
                                                  int ntail, unsigned long long* __restrict__ out) {
  unsigned long long acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)words[i] * (2ull * (i % 1000003ull) + 1ull);
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) acc += (unsigned long long)tail[threadIdx.x] * (977ull + threadIdx.x);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
int vrk_checksum(vr_ctx* ctx, const void* dev, size_t bytes, uint64_t* out) {
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->scratch) + 64;  // bytes 512.. of the 4 KiB scratch
  VR_CUDA(cudaMemsetAsync(acc, 0, 8, ctx->stream));
  const size_t nwords = bytes / 4;
  const unsigned blocks = (unsigned)std::min<size_t>(std::max<size_t>(div_up(nwords, 256), 1), (size_t)ctx->sm_count * 8);
  k_checksum<<<blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(dev), nwords, reinterpret_cast<const uint8_t*>(dev) + nwords * 4,
                                              (int)(bytes - nwords * 4), acc);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  unsigned long long* pin = reinterpret_cast<unsigned long long*>(ctx->scratch_host) + 64;
  VR_CUDA(cudaMemcpyAsync(pin, acc, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  *out = *pin;
  return VR_OK;
}

__global__ void __launch_bounds__(256) k_cache_gather(const uint2* __restrict__ cache, const uint32_t* __restrict__ idx, size_t n,
                                                      uint2* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cache[idx[i]];
}
int vrk_cache_gather(vr_ctx* ctx, const uint32_t* cache, const uint32_t* idx_dev, size_t n, uint2* out_dev) {
  k_cache_gather<<<div_up(n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint2*>(cache), idx_dev, n, out_dev);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
