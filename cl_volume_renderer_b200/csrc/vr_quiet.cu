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
// The step field holds, per voxel cell, 16 bits in a 3-D surface: low byte = the SDF value (what march() reads at trunc(origin),
// utility_ray.cl:148-150), high byte = the 8 verdicts (bit ux + 2 uy + 4 uz set = quiet).  Algorithmic bytes: 2 N (volume) + N (SDF)
// read, 2 N written.
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

// Round 2's first kernel computed, per voxel cell, the eight octant intervals from a three-plane window (one thread per cell
// column, ~120 instructions per cell, 0.96 ms at 512^3).  But the interval of octant (ux, uy, uz) of cell (x, y, z) is the min / max
// of the 2x2x2 texel block whose upper corner is c = (x + ux, y + uy, z + uz) — it belongs to the LATTICE CORNER c, and eight cells
// share it.  So:
//   k_lin_corners: Q(c) = 1 iff [min, max] of the texels {cx-1, cx} x {cy-1, cy} x {cz-1, cz} (0 outside the volume) meets no
//                  clause — one test per corner, one BIT per corner (a warp = 32 consecutive cx of a corner row, ballot = word),
//                  walking cz with the previous plane's 2x2 min / max in registers: two 2-byte loads per corner
//   k_lin_cells  : a thread assembles four cells of a row: 5 bits of each of the four corner rows (y, y+1) x (z, z+1) give the
//                  four verdict bytes (bit ux + 2 uy + 4 uz of cell i = Q(x + i + ux, y + uy, z + uz)), one 4-byte load the SDF
//                  bytes, one 8-byte surface write the four 16-bit entries
// spread the low 4 bits of b to 4 bytes 0x00/0x01
__device__ __forceinline__ uint32_t bits4_to_bytes4(uint32_t b) { return ((b & 0xFu) * 0x00204081u) & 0x01010101u; }
#define QC_Z 32  // corner planes per warp item
template <int NCL>  // clauses of the transfer function kept in registers (1 or 2: what the UI and the tests generate); 0 = any number
__global__ void __launch_bounds__(256) k_lin_corners(VolView vol, TfTable tf, uint32_t* __restrict__ Q, int qw, unsigned items, unsigned rows,
                                                     unsigned zchunks) {
  __shared__ TfIntervals iv;
  tf_intervals_init(tf, &iv, (int)threadIdx.x);
  __syncthreads();
  const int lo0 = iv.lo[0], hi0 = iv.hi[0], lo1 = NCL == 2 ? iv.lo[1] : 0, hi1 = NCL == 2 ? iv.hi[1] : 0;
  const unsigned lane = threadIdx.x & 31;
  const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
  const size_t nxy = (size_t)vol.nx * vol.ny;
  const size_t qplane = (size_t)rows * (unsigned)qw;
  for (unsigned it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += nwarps) {
    const unsigned xb = it % (unsigned)qw, t = it / (unsigned)qw;
    const int cy = (int)(t % rows), zc = (int)(t / rows);
    const int cx = (int)(xb * 32u + lane);
    const int z0 = zc * QC_Z, z1 = min(z0 + QC_Z, vol.nz + 1);
    // 2x2 min / max of texel plane z at this corner's (x, y): texels (cx-1, cx) x (cy-1, cy); texels outside the volume are 0.
    // Which of a lane's four loads exist is decided once per item (rows cy-1 / cy inside the volume, texel cx inside, lane 0 also
    // loads the texel left of the warp's segment; the other lanes get theirs by shuffle); pointers advance by a plane per step.
    const bool xin = cx < vol.nx, lin = lane == 0 && cx >= 1 && cx - 1 < vol.nx;
    const bool pa = cy >= 1 && xin, pb = cy < vol.ny && xin, pa0 = cy >= 1 && lin, pb0 = cy < vol.ny && lin;
    struct Tex4 { int a, b, a0, b0; };
    // texel (cx, cy - 1, z); the pointer may lie outside the allocation while no predicate lets it be used
    const int16_t* p = vol.v + ((ptrdiff_t)(z0 - 1) * (ptrdiff_t)nxy + (ptrdiff_t)(cy - 1) * vol.nx + cx);
    int zt = z0 - 1;  // texel plane p points into
    auto fetch = [&]() {
      const bool zv = (unsigned)zt < (unsigned)vol.nz;
      Tex4 q;
      q.a = (zv && pa) ? (int)__ldg(p) : 0;
      q.b = (zv && pb) ? (int)__ldg(p + vol.nx) : 0;
      q.a0 = (zv && pa0) ? (int)__ldg(p - 1) : 0;
      q.b0 = (zv && pb0) ? (int)__ldg(p + vol.nx - 1) : 0;
      p += nxy; ++zt;
      return q;
    };
    auto plane = [&](const Tex4& q, int* mn, int* mx) {
      int al = __shfl_up_sync(0xffffffffu, q.a, 1), bl = __shfl_up_sync(0xffffffffu, q.b, 1);
      if (lane == 0) { al = q.a0; bl = q.b0; }
      *mn = min(min(q.a, q.b), min(al, bl));
      *mx = max(max(q.a, q.b), max(al, bl));
    };
    int pmn, pmx;
    plane(fetch(), &pmn, &pmx);
    Tex4 cur = fetch(), nxt = fetch();  // two planes in flight: an iteration is far shorter than a load
    uint32_t* qout = Q + ((size_t)z0 * qplane + (size_t)cy * (unsigned)qw + xb);
    for (int cz = z0; cz < z1; ++cz, qout += qplane) {
      const Tex4 nn = fetch();
      int mn, mx;
      plane(cur, &mn, &mx);
      const int bmn = min(pmn, mn), bmx = max(pmx, mx);
      bool hit;  // the interval meets a clause
      if (NCL == 1) hit = (bmx >= lo0) & (bmn <= hi0);
      else if (NCL == 2) hit = ((bmx >= lo0) & (bmn <= hi0)) | ((bmx >= lo1) & (bmn <= hi1));
      else {
        hit = false;  // a uniform loop, no early exit: lanes do not diverge
        for (int i = 0; i < tf.n; ++i) hit |= (bmx >= iv.lo[i]) & (bmn <= iv.hi[i]);
      }
      const unsigned word = __ballot_sync(0xffffffffu, !hit && cx <= vol.nx);
      if (lane == 0) *qout = word;
      pmn = mn; pmx = mx;
      cur = nxt; nxt = nn;
    }
  }
}

__global__ void __launch_bounds__(256) k_lin_cells(SdfView sdf, const uint32_t* __restrict__ Q, int qw, unsigned rows, cudaSurfaceObject_t out,
                                                   unsigned groups_x, size_t ngroups) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += (size_t)gridDim.x * blockDim.x) {
    const int x0 = (int)(i % groups_x) * 4;
    const size_t t = i / groups_x;
    const int y = (int)(t % (unsigned)sdf.ny), z = (int)(t / (unsigned)sdf.ny);
    const unsigned wi = (unsigned)x0 >> 5, sh = (unsigned)x0 & 31u;
    uint32_t m = 0;  // byte i = the eight verdicts of cell x0 + i
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint32_t* row = Q + ((size_t)(z + (r >> 1)) * rows + (unsigned)(y + (r & 1))) * (unsigned)qw + wi;
      const uint32_t w0 = __ldg(row), w1 = (sh == 28u && (int)wi + 1 < qw) ? __ldg(row + 1) : 0u;
      const uint32_t b = __funnelshift_r(w0, w1, sh) & 0x1Fu;  // corners x0 .. x0 + 4 of this corner row
      m |= bits4_to_bytes4(b) << (2 * (r & 1) + 4 * (r >> 1));           // ux = 0
      m |= bits4_to_bytes4(b >> 1) << (1 + 2 * (r & 1) + 4 * (r >> 1));  // ux = 1
    }
    if (x0 + 4 <= sdf.nx) {
      const uint32_t d = __ldg(reinterpret_cast<const uint32_t*>(sdf.f + sdf.addr(x0, y, z)));  // four bytes of one brick row
      surf3Dwrite(make_uint2(__byte_perm(d, m, 0x5140), __byte_perm(d, m, 0x7362)), out, x0 * 2, y, z);
    } else {
      for (int k = 0; x0 + k < sdf.nx; ++k) {
        const unsigned d = (unsigned)(unsigned char)__ldg(sdf.f + sdf.addr(x0 + k, y, z));
        surf3Dwrite((unsigned short)((((m >> (8 * k)) & 0xFFu) << 8) | d), out, (x0 + k) * 2, y, z);
      }
    }
  }
}

int vrk_lin_field_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const int8_t* sdf_bricked, const TfTable& tf,
                        cudaSurfaceObject_t out) {
  VolView v{vol, nx, ny, nz};
  SdfView s{sdf_bricked, nx, ny, nz, nx / 8 + 1, ny / 8 + 1};
  const int qw = (nx + 1 + 31) / 32;
  const unsigned rows = (unsigned)ny + 1u, zchunks = (unsigned)div_up(nz + 1, QC_Z);
  uint32_t* Q = nullptr;
  VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&Q), (size_t)qw * rows * ((size_t)nz + 1) * sizeof(uint32_t), ctx->stream));
  const unsigned items = (unsigned)qw * rows * zchunks;
  const unsigned cgrid = (unsigned)std::min<size_t>(div_up(items, 8), (size_t)ctx->sm_count * 32);
  if (tf.n == 1) k_lin_corners<1><<<cgrid, 256, 0, ctx->stream>>>(v, tf, Q, qw, items, rows, zchunks);
  else if (tf.n == 2) k_lin_corners<2><<<cgrid, 256, 0, ctx->stream>>>(v, tf, Q, qw, items, rows, zchunks);
  else k_lin_corners<0><<<cgrid, 256, 0, ctx->stream>>>(v, tf, Q, qw, items, rows, zchunks);
  const unsigned groups_x = (unsigned)div_up(nx, 4);
  const size_t ngroups = (size_t)groups_x * ny * nz;
  k_lin_cells<<<(unsigned)std::min<size_t>(div_up(ngroups, 256), (size_t)ctx->sm_count * 32), 256, 0, ctx->stream>>>(s, Q, qw, rows, out, groups_x,
                                                                                                                     ngroups);
  ctx->launches += 2;
  VR_CUDA(cudaGetLastError());
  VR_CUDA(cudaFreeAsync(Q, ctx->stream));
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
