// vr_sdf.cu — signed distance field build.
//
// Reference: opencl_kernels/signed_distance_field.cl (create_base_image :6-54, neightbour_distance_calc :56-87,
// create_signed_distance_field :89-112) driven by app/signed_distance_field.cpp:7-35 — up to 129 full-volume
// ping-pong passes, each with two blocking 4-byte PCIe transfers.
//
// What the reference computes (DESIGN.md §4.2 has the proof sketch; tests/test_sdf_* pin it bit-exactly):
//   * event(v)  = is_event_gen(volume[v], |grad(v)|)
//   * base(v)   = s(v) * 1 on the event boundary band, s(v) * max_it where all 8 clamped corner neighbours
//                 share v's event state;  s(v) = -1 inside an event, +1 outside
//   * level i   : a voxel still at max_it whose smallest corner magnitude equals i becomes i+1 (sign kept)
//   i.e. a level-synchronous BFS over the corner-neighbour graph, capped at max_it.
// A value written during level i is i+1 > i, so it can neither satisfy nor break another voxel's `min == i`
// test in the same level: the update is hazard-free IN PLACE.  So this build keeps ONE int8 field (no
// ping-pong), never touches the host inside the loop, and stops launching work once a level finalises nothing.
#include "vr_device.cuh"

#define TX 32
#define TY 4
#define TZ 4

// pass 1: event bit per voxel (1 byte), evaluated once instead of 9 times per voxel as in create_base_image
__global__ void __launch_bounds__(TX* TY* TZ) k_sdf_event(VolView vol, TfTable tf, uint8_t* __restrict__ ev) {
  const int x = blockIdx.x * TX + threadIdx.x;
  const int y = blockIdx.y * TY + threadIdx.y;
  const int z = blockIdx.z * TZ + threadIdx.z;
  if (x >= vol.nx || y >= vol.ny || z >= vol.nz) return;
  ev[(size_t)x + (size_t)vol.nx * ((size_t)y + (size_t)vol.ny * (size_t)z)] = voxel_event(vol, tf, x, y, z) != 0;
}

// pass 2: create_base_image, signed_distance_field.cl:22-53
__global__ void __launch_bounds__(TX* TY* TZ) k_sdf_base(const uint8_t* __restrict__ ev, int nx, int ny, int nz,
                                                         int max_it, int8_t* __restrict__ field,
                                                         unsigned* __restrict__ level_count) {
  const int x = blockIdx.x * TX + threadIdx.x;
  const int y = blockIdx.y * TY + threadIdx.y;
  const int z = blockIdx.z * TZ + threadIdx.z;
  if (x >= nx || y >= ny || z >= nz) return;
  const size_t i = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z);
  const int e = ev[i];
  const int xm = max(x - 1, 0), xp = min(x + 1, nx - 1);
  const int ym = max(y - 1, 0), yp = min(y + 1, ny - 1);
  const int zm = max(z - 1, 0), zp = min(z + 1, nz - 1);
  bool homog = true;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int cx = (c & 1) ? xp : xm, cy = (c & 2) ? yp : ym, cz = (c & 4) ? zp : zm;
    homog &= (ev[(size_t)cx + (size_t)nx * ((size_t)cy + (size_t)ny * (size_t)cz)] == e);
  }
  int v = e ? -1 : 1;
  if (homog) v *= max_it;
  field[i] = (int8_t)v;
  if (!homog || max_it == 1) atomicAdd(level_count, 1u);  // band voxels seed level 1 (coarse: only != 0 matters)
}

// one BFS level, in place: create_signed_distance_field, signed_distance_field.cl:89-112
__global__ void __launch_bounds__(TX* TY* TZ) k_sdf_level(int8_t* __restrict__ field, int nx, int ny, int nz,
                                                          int iteration, int max_it,
                                                          const unsigned* __restrict__ prev_count,
                                                          unsigned* __restrict__ this_count) {
  if (*prev_count == 0) return;  // the previous level finalised nothing: the wavefront is dead
  const int x = blockIdx.x * TX + threadIdx.x;
  const int y = blockIdx.y * TY + threadIdx.y;
  const int z = blockIdx.z * TZ + threadIdx.z;
  bool changed = false;
  if (x < nx && y < ny && z < nz) {
    const size_t i = (size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z);
    const int local = field[i];
    const int absv = abs(local);
    if (absv > iteration) {
      const int xm = max(x - 1, 0), xp = min(x + 1, nx - 1);
      const int ym = max(y - 1, 0), yp = min(y + 1, ny - 1);
      const int zm = max(z - 1, 0), zp = min(z + 1, nz - 1);
      int nd = 127, abs_added = 0, added = 0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int cx = (c & 1) ? xp : xm, cy = (c & 2) ? yp : ym, cz = (c & 4) ? zp : zm;
        const int val = field[(size_t)cx + (size_t)nx * ((size_t)cy + (size_t)ny * (size_t)cz)];
        const int a = abs(val);
        abs_added += a;
        added += val;
        nd = min(nd, a);
      }
      if (abs(added) != abs_added) nd = 0;
      if (nd != 0 && nd == iteration && iteration + 1 < max_it) {
        field[i] = (int8_t)(local < 0 ? -(iteration + 1) : (iteration + 1));
        changed = true;
      }
    }
  }
  const unsigned any = __ballot_sync(0xffffffffu, changed);
  if (any && (threadIdx.x & 31) == 0) atomicAdd(this_count, (unsigned)__popc(any));
}

int vrk_sdf_build(vr_ctx* ctx, const int16_t* vol, int nx, int ny, int nz, const TfTable& tf, int8_t* field,
                  int* levels_out, int* max_it_out) {
  const size_t n = (size_t)nx * ny * nz;
  const int max_it = std::min(std::max(nx, std::max(ny, nz)) / 2, 127);  // signed_distance_field.cpp:11
  uint8_t* ev = nullptr;
  unsigned* counts = nullptr;
  VR_CUDA(cudaMallocAsync(&ev, n, ctx->stream));
  VR_CUDA(cudaMallocAsync(&counts, sizeof(unsigned) * 130, ctx->stream));
  VR_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned) * 130, ctx->stream));
  VolView v{vol, nx, ny, nz};
  dim3 grid(div_up(nx, TX), div_up(ny, TY), div_up(nz, TZ)), block(TX, TY, TZ);
  k_sdf_event<<<grid, block, 0, ctx->stream>>>(v, tf, ev);
  k_sdf_base<<<grid, block, 0, ctx->stream>>>(ev, nx, ny, nz, max_it, field, counts + 0);
  ctx->launches += 2;
  // level i finalises magnitude i+1, which is only stored when i+1 < max_it
  for (int it = 1; it + 1 < max_it; ++it) {
    k_sdf_level<<<grid, block, 0, ctx->stream>>>(field, nx, ny, nz, it, max_it, counts + it - 1, counts + it);
    ctx->launches++;
  }
  VR_CUDA(cudaGetLastError());
  unsigned* hc = reinterpret_cast<unsigned*>(ctx->scratch_host);
  VR_CUDA(cudaMemcpyAsync(hc, counts, sizeof(unsigned) * 130, cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaFreeAsync(ev, ctx->stream));
  VR_CUDA(cudaFreeAsync(counts, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  int levels = 0;
  for (int it = 1; it + 1 < max_it; ++it)
    if (hc[it] != 0) levels = it;
  *levels_out = levels;
  *max_it_out = max_it;
  return VR_OK;
}
