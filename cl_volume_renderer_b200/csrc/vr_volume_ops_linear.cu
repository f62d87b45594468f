// vr_volume_ops_linear.cu — fetch_stats, tf_sort_values and bilateral_filter under VR_SAMPLING_HW_LINEAR: the reading NVIDIA hardware
// gives the reference's kernels AS SHIPPED (every sampler asks for CLK_FILTER_LINEAR on the int16 volume).
//
// The texel values come from the texture unit itself (normalised-float read of an int16 CUDA array, linear filter):
// rint(t * 32767) equals the OpenCL runtime's read_imagei bit for bit (profiles/r1b_cuda_texture_vs_opencl_linear.txt).  Two texture
// objects over the same array: `border` = the CLK_ADDRESS_CLAMP sampler of gradient_prewitt_nn / bilateral_kernel
// (utility_filter.cl:4,40), `edge` = the sampler of fetch_stats / tf_sort_values that names no addressing mode
// (reference_volume_figures.cl:12, histogram.cl:7), which the hardware serves like clamp-to-edge (fitted against the recorded
// runs, tests/golden/opencl_reference_runs.npz).  All arithmetic after the fetch is the NEAREST kernels' (vr_volume_ops.cu).
// Opt-in and simple: one thread per voxel, no tiling — every tap is a hardware-filtered fetch of its own.
#include <cstring>

#include "vr_device.cuh"

#define LX 32
#define LY 4
#define LZ 4

__device__ __forceinline__ int tex_int16(cudaTextureObject_t t, float x, float y, float z) {
  return __double2int_rn((double)tex3D<float>(t, x, y, z) * 32767.0);
}
__device__ __forceinline__ f3 gradient_tex(cudaTextureObject_t border, float x, float y, float z) {  // utility_filter.cl:2-35
  const int dx = tex_int16(border, x + 1.0f, y, z) - tex_int16(border, x - 1.0f, y, z);
  const int dy = tex_int16(border, x, y + 1.0f, z) - tex_int16(border, x, y - 1.0f, z);
  const int dz = tex_int16(border, x, y, z + 1.0f) - tex_int16(border, x, y, z - 1.0f);
  return {(float)dx, (float)dy, (float)dz};
}

// fetch_stats, reference_volume_figures.cl:10-26
__global__ void __launch_bounds__(LX* LY* LZ) k_fetch_stats_linear(cudaTextureObject_t border, cudaTextureObject_t edge, int nx, int ny,
                                                                   int nz, int32_t* __restrict__ stats, int zlo, int zhi) {
  const int x = blockIdx.x * LX + threadIdx.x, y = blockIdx.y * LY + threadIdx.y, z = blockIdx.z * LZ + threadIdx.z;
  int mnv = INT32_MAX, mxv = INT32_MIN, mng = INT32_MAX, mxg = INT32_MIN;
  if (x < nx && y < ny && z >= zlo && z < zhi && z < nz) {
    const int v = tex_int16(edge, (float)x, (float)y, (float)z);
    const int g = f2i(length3(gradient_tex(border, (float)x, (float)y, (float)z)));
    mnv = mxv = v;
    mng = mxg = g;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnv = min(mnv, __shfl_xor_sync(0xffffffffu, mnv, o));
    mxv = max(mxv, __shfl_xor_sync(0xffffffffu, mxv, o));
    mng = min(mng, __shfl_xor_sync(0xffffffffu, mng, o));
    mxg = max(mxg, __shfl_xor_sync(0xffffffffu, mxg, o));
  }
  if (threadIdx.x == 0) {  // one warp per (y, z) row of the block
    atomicMin(stats + 0, mnv); atomicMax(stats + 1, mxv);
    atomicMin(stats + 2, mng); atomicMax(stats + 3, mxg);
  }
}

// tf_sort_values, histogram.cl:4-32 (out-of-range indices dropped, y == height aliasing kept: SURVEY A.5)
__global__ void __launch_bounds__(LX* LY* LZ) k_histogram_linear(cudaTextureObject_t border, cudaTextureObject_t edge, int nx, int ny, int nz,
                                                                 uint32_t* __restrict__ bins, int width, int height, float min_v,
                                                                 float max_v, float min_g, float max_g, int zlo, int zhi) {
  const int x = blockIdx.x * LX + threadIdx.x, y = blockIdx.y * LY + threadIdx.y, z = blockIdx.z * LZ + threadIdx.z;
  long long flat = -1;
  if (x < nx && y < ny && z >= zlo && z < zhi && z < nz) {
    const int ref_value = tex_int16(edge, (float)x, (float)y, (float)z);
    const float g = length3(gradient_tex(border, (float)x, (float)y, (float)z));
    if (!(g > max_g) && !((float)ref_value > max_v)) {
      const float value_range = max_v - min_v, gradient_range = max_g - min_g;
      const int px = f2i(roundf((((float)ref_value - min_v) / value_range) * (float)width));
      const int py = f2i(roundf(((g - min_g) / gradient_range) * (float)height));
      flat = (long long)px * height + py;
      if (flat < 0 || flat >= (long long)width * height) flat = -1;
    }
  }
  const unsigned active = __ballot_sync(0xffffffffu, flat >= 0);
  if (flat >= 0) {
    const unsigned peers = __match_any_sync(active, (int)flat);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(bins + flat, (uint32_t)__popc(peers));
  }
}

// bilateral_filter, volume_filter.cl:5-11 + utility_filter.cl:38-62: 125 filtered taps, pow(d, 2) as d*d
__global__ void __launch_bounds__(LX* LY* LZ) k_bilateral_linear(cudaTextureObject_t border, int nx, int ny, int nz, int16_t* __restrict__ dst) {
  const int x = blockIdx.x * LX + threadIdx.x, y = blockIdx.y * LY + threadIdx.y, z = blockIdx.z * LZ + threadIdx.z;
  if (x >= nx || y >= ny || z >= nz) return;
  const float sigmas = 0.6f, sigmar = 1.0f;
  const float mid = (float)tex_int16(border, (float)x, (float)y, (float)z);
  float out_colour = 0.0f, wp = 0.0f;
  for (int dz = -2; dz <= 2; ++dz)
    for (int dy = -2; dy <= 2; ++dy)
      for (int dx = -2; dx <= 2; ++dx) {
        const float local = (float)tex_int16(border, (float)x + (float)dx, (float)y + (float)dy, (float)z + (float)dz);
        const float posd = ((float)(dx * dx + dy * dy + dz * dz)) / (2 * sigmas * sigmas);
        const float diff = mid - local;
        const float cold = (diff * diff) / (2 * sigmar * sigmar);
        const float w = expf(-posd - cold);
        wp += w;
        out_colour += local * w;
      }
  dst[(size_t)x + (size_t)nx * ((size_t)y + (size_t)ny * (size_t)z)] = (int16_t)f2s(out_colour / wp);
}

int vrk_fetch_stats_linear(vr_ctx* ctx, cudaTextureObject_t border, cudaTextureObject_t edge, int nx, int ny, int nz, int32_t out[4],
                           int zlo, int zhi) {
  const int32_t init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
  memcpy(ctx->scratch_host, init, sizeof(init));
  VR_CUDA(cudaMemcpyAsync(ctx->scratch, ctx->scratch_host, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  dim3 grid(div_up(nx, LX), div_up(ny, LY), div_up(nz, LZ)), block(LX, LY, LZ);
  k_fetch_stats_linear<<<grid, block, 0, ctx->stream>>>(border, edge, nx, ny, nz, ctx->scratch, zlo, zhi);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  VR_CUDA(cudaMemcpyAsync(ctx->scratch_host, ctx->scratch, sizeof(init), cudaMemcpyDeviceToHost, ctx->stream));
  VR_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->scratch_host, sizeof(init));
  return VR_OK;
}

int vrk_histogram_linear(vr_ctx* ctx, cudaTextureObject_t border, cudaTextureObject_t edge, int nx, int ny, int nz, int width, int height,
                         const float range[4], uint32_t* bins_dev, int zlo, int zhi) {
  VR_CUDA(cudaMemsetAsync(bins_dev, 0, sizeof(uint32_t) * (size_t)width * height, ctx->stream));
  dim3 grid(div_up(nx, LX), div_up(ny, LY), div_up(nz, LZ)), block(LX, LY, LZ);
  k_histogram_linear<<<grid, block, 0, ctx->stream>>>(border, edge, nx, ny, nz, bins_dev, width, height, range[0], range[1], range[2],
                                                      range[3], zlo, zhi);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

int vrk_bilateral_linear(vr_ctx* ctx, cudaTextureObject_t border, int nx, int ny, int nz, int16_t* dst) {
  dim3 grid(div_up(nx, LX), div_up(ny, LY), div_up(nz, LZ)), block(LX, LY, LZ);
  k_bilateral_linear<<<grid, block, 0, ctx->stream>>>(border, nx, ny, nz, dst);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
