// vr_quiet.cu — the step field of VR_SAMPLING_HW_LINEAR (no counterpart in the reference; see vr_render.cu `lin_cell`).
//
// With the reference's samplers as NVIDIA hardware executes them (CLK_FILTER_LINEAR on the int16 volume, utility_ray.cl:130-132,
// utility_filter.cl:4) the value get_event_and_value tests at a position p is
//     R( sum_w w * texel ) ,  eight weights >= 0 that sum to 256/256, texels = the 2x2x2 block at the hardware cell
//     c = floor((floor(p*256 + 1/2) - 128) / 256) per axis  (oracle.cpp hw_linear_fetch, pinned on 874 545 probe samples)
// i.e. an integer inside [min, max] of those eight texels (border texels read 0).  If that interval meets no clause of the
// transfer function — `value >= min_v && value <= max_v`, or `value > K` (tf_part.cpp:60-77) — no event can occur at p, whatever
// the gradient clause says.  c is floor(p) - 1 or floor(p), so every voxel cell floor(p) has 8 octants with one verdict each.
//
// k_lin_field writes, per voxel cell, 16 bits into a 3-D surface: low byte = the SDF value (what march() reads at trunc(origin),
// utility_ray.cl:148-150), high byte = the 8 verdicts (bit ux + 2 uy + 4 uz set = quiet).  A thread owns an (x, y) column of ZC
// cells and walks z with a three-plane window of per-plane quadrant intervals; a warp is 32 consecutive x of one row, so per
// plane a lane loads 3 texels and gets its x neighbours by shuffle.  One test of the whole neighbourhood's interval settles all
// eight octants of most cells.  (First version, 9 loads per plane and float compares per octant: 351 instructions per cell,
// 1.47 ms at 512^3, issue-bound.)  Algorithmic bytes: 2 N (volume) + N (SDF) read, 2 N written.
#include "vr_device.cuh"

// The clauses as integer intervals: voxel values are integers, so `(float)v >= min_v` is `v >= ceil(min_v)`, `(float)v <= max_v` is
// `v <= floor(max_v)` and the threshold form `(float)v > K` is `v >= floor(K) + 1` — evaluated once per block into shared memory,
// so the per-cell tests are integer compares.  An interval [mn, mx] can meet clause i iff mx >= lo[i] && mn <= hi[i].
struct TfIntervals {
  int lo[VR_TF_MAX_RECTS], hi[VR_TF_MAX_RECTS];
};
__device__ __forceinline__ void tf_intervals_init(const TfTable& tf, TfIntervals* s, int tid) {
  if (tid < tf.n) {
    const vr_tf_rect& q = tf.r[tid];
    const float big = 100000.0f;  // beyond any int16 value
    int lo, hi;
    if (q.flags & VR_TF_THRESHOLD) {
      lo = (q.min_v == q.min_v) ? (int)floorf(fminf(fmaxf(q.min_v, -big), big)) + 1 : INT32_MAX;
      hi = INT32_MAX;
    } else {
      lo = (q.min_v == q.min_v) ? (int)ceilf(fminf(fmaxf(q.min_v, -big), big)) : INT32_MAX;   // NaN bound: the clause never matches
      hi = (q.max_v == q.max_v) ? (int)floorf(fminf(fmaxf(q.max_v, -big), big)) : INT32_MIN;
    }
    s->lo[tid] = lo; s->hi[tid] = hi;
  }
}
__device__ __forceinline__ bool tf_interval_quiet(const TfIntervals& s, int n, int mn, int mx) {
  for (int i = 0; i < n; ++i)
    if (mx >= s.lo[i] && mn <= s.hi[i]) return false;
  return true;
}

struct Quad {
  int lo[4], hi[4];  // [ux + 2*uy]: interval of the texels {x-1+ux, x+ux} x {y-1+uy, y+uy} of one plane
};

// A warp owns 32 consecutive x of one row: a lane loads its own texel of the rows y-1, y, y+1, the x neighbours come from the
// neighbouring lanes (lanes 0 and 1 also load the two texels beyond the ends of the warp's segment).
__device__ __forceinline__ Quad plane_quadrants(const VolView& vol, int x0, unsigned lane, int y, int z) {
  int t[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y - 1 + dy;
    const int v = vol.at(x0 + (int)lane, yy, z);  // outside the volume: border colour 0
    const int h = lane < 2 ? vol.at(lane == 0 ? x0 - 1 : x0 + 32, yy, z) : 0;
    int l = __shfl_up_sync(0xffffffffu, v, 1), r = __shfl_down_sync(0xffffffffu, v, 1);
    const int h0 = __shfl_sync(0xffffffffu, h, 0), h1 = __shfl_sync(0xffffffffu, h, 1);
    if (lane == 0) l = h0;
    if (lane == 31) r = h1;
    t[dy][0] = l; t[dy][1] = v; t[dy][2] = r;
  }
  int alo[3][2], ahi[3][2];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    alo[dy][0] = min(t[dy][0], t[dy][1]); ahi[dy][0] = max(t[dy][0], t[dy][1]);
    alo[dy][1] = min(t[dy][1], t[dy][2]); ahi[dy][1] = max(t[dy][1], t[dy][2]);
  }
  Quad q;
#pragma unroll
  for (int uy = 0; uy < 2; ++uy)
#pragma unroll
    for (int ux = 0; ux < 2; ++ux) {
      q.lo[ux + 2 * uy] = min(alo[uy][ux], alo[uy + 1][ux]);
      q.hi[ux + 2 * uy] = max(ahi[uy][ux], ahi[uy + 1][ux]);
    }
  return q;
}

template <int ZC>
__global__ void __launch_bounds__(256) k_lin_field(VolView vol, SdfView sdf, TfTable tf, cudaSurfaceObject_t out) {
  __shared__ TfIntervals iv;
  tf_intervals_init(tf, &iv, (int)threadIdx.x);
  __syncthreads();
  const unsigned lane = threadIdx.x & 31;
  const int x0 = blockIdx.x * 32, x = x0 + (int)lane;
  const int y = blockIdx.y * 8 + (int)(threadIdx.x >> 5);  // one row per warp
  const int z0 = blockIdx.z * ZC;
  if (y >= vol.ny) return;
  Quad m = plane_quadrants(vol, x0, lane, y, z0 - 1), c = plane_quadrants(vol, x0, lane, y, z0);
  const int z1 = min(z0 + ZC, vol.nz);
  for (int z = z0; z < z1; ++z) {
    const Quad n = plane_quadrants(vol, x0, lane, y, z + 1);
    unsigned mask = 0xFFu;
    // most cells are far from any surface: one test of the whole 3x3x3 neighbourhood settles all eight octants
    int ulo = min(min(m.lo[0], m.lo[1]), min(m.lo[2], m.lo[3])), uhi = max(max(m.hi[0], m.hi[1]), max(m.hi[2], m.hi[3]));
    ulo = min(ulo, min(min(c.lo[0], c.lo[1]), min(c.lo[2], c.lo[3]))); uhi = max(uhi, max(max(c.hi[0], c.hi[1]), max(c.hi[2], c.hi[3])));
    ulo = min(ulo, min(min(n.lo[0], n.lo[1]), min(n.lo[2], n.lo[3]))); uhi = max(uhi, max(max(n.hi[0], n.hi[1]), max(n.hi[2], n.hi[3])));
    if (!tf_interval_quiet(iv, tf.n, ulo, uhi)) {
      mask = 0;
#pragma unroll
      for (int oct = 0; oct < 8; ++oct) {
        const int q = oct & 3;
        const int mn = (oct & 4) ? min(c.lo[q], n.lo[q]) : min(m.lo[q], c.lo[q]);
        const int mx = (oct & 4) ? max(c.hi[q], n.hi[q]) : max(m.hi[q], c.hi[q]);
        if (tf_interval_quiet(iv, tf.n, mn, mx)) mask |= 1u << oct;
      }
    }
    if (x < vol.nx) {
      const unsigned d = (unsigned)(unsigned char)__ldg(sdf.f + sdf.addr(x, y, z));
      surf3Dwrite((unsigned short)((mask << 8) | d), out, x * 2, y, z);
    }
    m = c;
    c = n;
  }
}

int vrk_lin_field_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const int8_t* sdf_bricked, const TfTable& tf,
                        cudaSurfaceObject_t out) {
  constexpr int ZC = 16;
  VolView v{vol, nx, ny, nz};
  SdfView s{sdf_bricked, nx, ny, nz, nx / 8 + 1, ny / 8 + 1};
  dim3 grid(div_up(nx, 32), div_up(ny, 8), div_up(nz, ZC));
  k_lin_field<ZC><<<grid, 256, 0, ctx->stream>>>(v, s, tf, out);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// the verdict bytes back as a linear array (tests: compared bit for bit with the oracle's orc_quiet_cells)
__global__ void __launch_bounds__(256) k_lin_masks(cudaSurfaceObject_t field, int nx, int ny, int nz, uint8_t* __restrict__ out) {
  const size_t n = (size_t)nx * ny * nz;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % nx);
    const size_t t = i / nx;
    const int y = (int)(t % ny), z = (int)(t / ny);
    out[i] = (uint8_t)(surf3Dread<unsigned short>(field, x * 2, y, z, cudaBoundaryModeZero) >> 8);
  }
}

int vrk_lin_field_masks(vr_ctx* ctx, cudaSurfaceObject_t field, int nx, int ny, int nz, uint8_t* masks_dev) {
  const size_t n = (size_t)nx * ny * nz;
  k_lin_masks<<<(unsigned)std::min<size_t>(div_up(n, 256), (size_t)ctx->sm_count * 16), 256, 0, ctx->stream>>>(field, nx, ny, nz, masks_dev);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
