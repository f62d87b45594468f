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
// cells and walks z with a three-plane window of per-plane quadrant intervals: 9 volume loads (L1-resident: neighbouring threads
// read the same rows) and 8 interval tests per cell.  Algorithmic bytes: 2 N (volume) + N (SDF) read, 2 N written.
#include "vr_device.cuh"

__device__ __forceinline__ bool tf_interval_quiet(const TfTable& tf, int mn, int mx) {
  for (int i = 0; i < tf.n; ++i) {
    const vr_tf_rect& q = tf.r[i];
    if (q.flags & VR_TF_THRESHOLD) {
      if ((float)mx > q.min_v) return false;
    } else if ((float)mx >= q.min_v && (float)mn <= q.max_v) {
      return false;
    }
  }
  return true;
}

struct Quad {
  int lo[4], hi[4];  // [ux + 2*uy]: interval of the texels {x-1+ux, x+ux} x {y-1+uy, y+uy} of one plane
};

__device__ __forceinline__ Quad plane_quadrants(const VolView& vol, int x, int y, int z) {
  int t[3][3];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) t[dy][dx] = vol.at(x - 1 + dx, y - 1 + dy, z);  // outside the volume: border colour 0
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
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int z0 = blockIdx.z * ZC;
  if (x >= vol.nx || y >= vol.ny) return;
  Quad m = plane_quadrants(vol, x, y, z0 - 1), c = plane_quadrants(vol, x, y, z0);
  const int z1 = min(z0 + ZC, vol.nz);
  for (int z = z0; z < z1; ++z) {
    const Quad n = plane_quadrants(vol, x, y, z + 1);
    unsigned mask = 0;
#pragma unroll
    for (int oct = 0; oct < 8; ++oct) {
      const int q = oct & 3;
      const int mn = (oct & 4) ? min(c.lo[q], n.lo[q]) : min(m.lo[q], c.lo[q]);
      const int mx = (oct & 4) ? max(c.hi[q], n.hi[q]) : max(m.hi[q], c.hi[q]);
      if (tf_interval_quiet(tf, mn, mx)) mask |= 1u << oct;
    }
    const unsigned d = (unsigned)(unsigned char)__ldg(sdf.f + sdf.addr(x, y, z));
    surf3Dwrite((unsigned short)((mask << 8) | d), out, x * 2, y, z);
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
