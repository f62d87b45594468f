// vr_frame_filter.cu — 2-D filter over an RGBA8 frame: opencl_kernels/2d_image_filter.cl:1-43.
//
// The reference ships this kernel ("Applies a 2D bilateral filter (not tested)") but no host code ever launches it.
// Two modes (include/vr.h):
//   VR_FILTER2D_REFERENCE  the kernel's arithmetic exactly as written — bit-identical to the reference source compiled for
//                          the CPU (oracle/_ref) and to the restatement in oracle/oracle.cpp, quirks included:
//                          gauss() is a quotient a/(2σ²) without an exponential (:1-3), the spatial term is taken from
//                          pos - off (:26), the colour term from an UNSIGNED difference (:27-29), the sums weigh the
//                          centre colour (:30-32), every channel is divided by the red weight sum (:39-41), alpha = 0.
//   VR_FILTER2D_BILATERAL  what it set out to be: exp(-d²/2σ²)·exp(-Δc²/2σ²) weights per channel, centre included,
//                          taps outside the frame skipped, rounded to nearest, alpha kept (our definition).
// The reference kernel is in place on a __read_write image, so its taps race with the neighbours' writes; both modes
// here read the unfiltered frame only (src != dst).
//
// Memory-bound on paper (8 B per pixel), in practice (2k+1)² taps per pixel served by shared memory: a block's
// (32+2k) x (8+2k) tile is staged once with coalesced 4-byte loads (k <= 8; larger kernels read through L1).
#include "vr_device.cuh"

#define FB_X 32
#define FB_Y 8
#define FB_TILE_K 8  // largest kernel_size served from the shared-memory tile

struct FrameView {
  const uchar4* p;
  int w, h;
  __device__ __forceinline__ uchar4 border0(int x, int y) const {  // CLK_ADDRESS_CLAMP: border colour (0,0,0,0)
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return make_uchar4(0, 0, 0, 0);
    return __ldg(p + (size_t)y * w + x);
  }
};

// stages the tile around the block; outside texels hold `pad`
template <bool TILED>
__device__ __forceinline__ void stage_tile(const FrameView& f, uchar4* tile, int k, int bx0, int by0) {
  if (!TILED) return;
  const int tw = FB_X + 2 * k, th = FB_Y + 2 * k;
  for (int i = threadIdx.y * FB_X + threadIdx.x; i < tw * th; i += FB_X * FB_Y) {
    const int ty = i / tw, tx = i - ty * tw;
    tile[i] = f.border0(bx0 - k + tx, by0 - k + ty);
  }
  __syncthreads();
}

template <bool TILED>
__global__ void __launch_bounds__(FB_X* FB_Y) k_filter2d_reference(FrameView f, uchar4* __restrict__ dst, int k, float den) {
  extern __shared__ uchar4 tile[];
  const int bx0 = blockIdx.x * FB_X, by0 = blockIdx.y * FB_Y;
  stage_tile<TILED>(f, tile, k, bx0, by0);
  const int px = bx0 + threadIdx.x, py = by0 + threadIdx.y;
  if (px >= f.w || py >= f.h) return;
  const int tw = FB_X + 2 * k;
  const uchar4 c = TILED ? tile[(threadIdx.y + k) * tw + threadIdx.x + k] : f.border0(px, py);
  const unsigned ref0 = c.x, ref1 = c.y, ref2 = c.z;
  const float fr0 = (float)ref0, fr1 = (float)ref1, fr2 = (float)ref2;
  float wpr = 0.0f, r = 0.0f, g = 0.0f, b = 0.0f;
  for (int x = -k; x <= k; ++x) {      // 2d_image_filter.cl:21-22: x is the outer loop
    const float fx = (float)(px - x);
    const float fx2 = fx * fx;         // pow(v, 2)
    for (int y = -k; y <= k; ++y) {
      if (x == 0 && y == 0) continue;
      const uchar4 t = TILED ? tile[(threadIdx.y + k + y) * tw + threadIdx.x + k + x] : f.border0(px + x, py + y);
      const float fy = (float)(py - y);
      const float g1 = (fx2 + fy * fy) / den;
      const float g2r = fabsf((float)((unsigned)t.x - ref0)) / den;  // unsigned wrap-around as in the source
      const float g2g = fabsf((float)((unsigned)t.y - ref1)) / den;
      const float g2b = fabsf((float)((unsigned)t.z - ref2)) / den;
      r += fr0 * g1 * g2r;
      g += fr1 * g1 * g2g;
      b += fr2 * g1 * g2b;
      wpr += g1 * g2r;  // Wpg / Wpb are accumulated by the source but never used (:39-41)
    }
  }
  uchar4 o;
  o.x = (unsigned char)min(f2u(r / wpr), 255u);  // float -> uint: toward zero, saturating, NaN -> 0; write_imageui saturates
  o.y = (unsigned char)min(f2u(g / wpr), 255u);
  o.z = (unsigned char)min(f2u(b / wpr), 255u);
  o.w = 0;
  dst[(size_t)py * f.w + px] = o;
}

// Corrected bilateral.  A spatial weight takes (2k+1)² values and a colour weight 256; both tables are evaluated once per
// block with the expressions a per-tap evaluation would use, so the result does not depend on the tabulation.
#define FB_MAX_K 15
template <bool TILED>
__global__ void __launch_bounds__(FB_X* FB_Y) k_filter2d_bilateral(FrameView f, uchar4* __restrict__ dst, int k, float den) {
  extern __shared__ uchar4 tile[];
  __shared__ float ws[(2 * FB_MAX_K + 1) * (2 * FB_MAX_K + 1)];
  __shared__ float wc[256];
  const int side = 2 * k + 1;
  const int tid = threadIdx.y * FB_X + threadIdx.x;
  for (int i = tid; i < side * side; i += FB_X * FB_Y) {
    const int y = i / side - k, x = i - (i / side) * side - k;
    ws[i] = expf(-((float)(x * x + y * y) / den));
  }
  for (int i = tid; i < 256; i += FB_X * FB_Y) {
    const float d = (float)i;
    wc[i] = expf(-((d * d) / den));
  }
  const int bx0 = blockIdx.x * FB_X, by0 = blockIdx.y * FB_Y;
  stage_tile<TILED>(f, tile, k, bx0, by0);
  if (!TILED) __syncthreads();
  const int px = bx0 + threadIdx.x, py = by0 + threadIdx.y;
  if (px >= f.w || py >= f.h) return;
  const int tw = FB_X + 2 * k;
  const uchar4 c = TILED ? tile[(threadIdx.y + k) * tw + threadIdx.x + k] : f.border0(px, py);
  float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
  for (int y = -k; y <= k; ++y) {
    if ((unsigned)(py + y) >= (unsigned)f.h) continue;
    for (int x = -k; x <= k; ++x) {
      if ((unsigned)(px + x) >= (unsigned)f.w) continue;
      const uchar4 t = TILED ? tile[(threadIdx.y + k + y) * tw + threadIdx.x + k + x] : f.border0(px + x, py + y);
      const float s = ws[(y + k) * side + x + k];
      const float u0 = s * wc[abs((int)t.x - (int)c.x)];
      const float u1 = s * wc[abs((int)t.y - (int)c.y)];
      const float u2 = s * wc[abs((int)t.z - (int)c.z)];
      a0 += (float)t.x * u0; w0 += u0;
      a1 += (float)t.y * u1; w1 += u1;
      a2 += (float)t.z * u2; w2 += u2;
    }
  }
  uchar4 o;
  o.x = (unsigned char)min(f2u(a0 / w0 + 0.5f), 255u);
  o.y = (unsigned char)min(f2u(a1 / w1 + 0.5f), 255u);
  o.z = (unsigned char)min(f2u(a2 / w2 + 0.5f), 255u);
  o.w = c.w;
  dst[(size_t)py * f.w + px] = o;
}

int vrk_filter2d(vr_ctx* ctx, const uchar4* src, uchar4* dst, int w, int h, int kernel_size, float sigma, int mode) {
  const float den = 2 * sigma * sigma;  // gauss(): a / (2*sigma*sigma), 2d_image_filter.cl:2 — (2*sigma)*sigma in fp32
  FrameView f{src, w, h};
  dim3 grid(div_up(w, FB_X), div_up(h, FB_Y)), block(FB_X, FB_Y);
  const bool tiled = kernel_size <= FB_TILE_K;
  const size_t smem = tiled ? sizeof(uchar4) * (FB_X + 2 * kernel_size) * (FB_Y + 2 * kernel_size) : 0;
  if (mode == VR_FILTER2D_REFERENCE) {
    if (tiled) k_filter2d_reference<true><<<grid, block, smem, ctx->stream>>>(f, dst, kernel_size, den);
    else k_filter2d_reference<false><<<grid, block, 0, ctx->stream>>>(f, dst, kernel_size, den);
  } else {
    if (tiled) k_filter2d_bilateral<true><<<grid, block, smem, ctx->stream>>>(f, dst, kernel_size, den);
    else k_filter2d_bilateral<false><<<grid, block, 0, ctx->stream>>>(f, dst, kernel_size, den);
  }
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
