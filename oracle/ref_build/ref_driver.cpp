// ref_driver.cpp — host side of oracle/_ref: runs the reference's own OpenCL kernels (compiled as C++ through
// cl_emu.hpp) over NDRanges on the CPU, with the host loops restated from the reference's app/*.cpp (which cannot be
// compiled here: they need <CL/opencl.h>).
//
// TEST INFRASTRUCTURE: used by tests/ to pin oracle/oracle.cpp against the reference's actual kernel source, and by
// bench.py --impl reference as the "reference kernels on the host cores" arm.  Never linked into the product.
//
// The gen_*.inc files are produced at build time by build_ref.py from /root/reference/opencl_kernels/*.cl (include
// expansion only, plus ONE syntactic fix documented there) and deleted after compilation.
#include "cl_emu.hpp"

#include <algorithm>
#include <cstring>
#include <set>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

thread_local emu_ndrange emu_wi;
thread_local emu_tf emu_tf_cur;

namespace k_ray_marching {
#include "gen_ray_marching.inc"
}
namespace k_sdf {
#include "gen_signed_distance_field.inc"
}
namespace k_histogram {
#include "gen_histogram.inc"
}
namespace k_volume_filter {
#include "gen_volume_filter.inc"
}
namespace k_buffer_reset {
#include "gen_buffer_reset.inc"
}
namespace k_figures {
#include "gen_reference_volume_figures.inc"
}
namespace k_clip {
#include "gen_reference_volume_clip.inc"
}
namespace k_image_filter {
#include "gen_2d_image_filter.inc"
}

namespace {
// app/common.hpp:59-66
unsigned evenness(unsigned g, unsigned l) { unsigned m = g % l; return m == 0 ? g : g + l - m; }

// clEnqueueNDRangeKernel over {gx,gy,gz}: one call of `f` per work-item
template <class F>
void ndrange(size_t gx, size_t gy, size_t gz, emu_tf tf, int threads, F f) {
  const long n_rows = (long)(gy * gz);
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads)
  for (long row = 0; row < n_rows; ++row) {
    emu_tf_cur = tf;
    const size_t y = (size_t)row % gy, z = (size_t)row / gy;
    for (size_t x = 0; x < gx; ++x) {
      emu_wi.gid[0] = x; emu_wi.gid[1] = y; emu_wi.gid[2] = z;
      f();
    }
  }
}
int nthreads(int t) {
#ifdef _OPENMP
  return t > 0 ? t : omp_get_max_threads();
#else
  return 1;
#endif
}
}  // namespace

extern "C" {

int ref_num_threads(void) { return nthreads(0); }

// app/signed_distance_field.cpp:7-35 around create_base_image / create_signed_distance_field
int ref_sdf_build(const int16_t* vol, int nx, int ny, int nz, const emu_tf_rect* rects, int n_rects, int8_t* out, int threads) {
  const size_t N = (size_t)nx * ny * nz;
  emu_tf tf{rects, n_rects};
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  std::vector<int8_t> a(N, 0), b(N, 0);
  emu_image sdf{a.data(), nx, ny, nz, EMU_S8}, sdf_pong{b.data(), nx, ny, nz, EMU_S8};
  size_t max_iterations = std::min((size_t)std::max(nx, std::max(ny, nz)) / 2, (size_t)127);
  const size_t gx = evenness(nx, 8), gy = evenness(ny, 8), gz = evenness(nz, 8);
  const int T = nthreads(threads);
  ndrange(gx, gy, gz, tf, T, [&] { k_sdf::create_base_image(&v, &sdf, &sdf_pong, (unsigned)max_iterations); });
  emu_image *ping = &sdf, *pong = &sdf_pong;
  int counter = 0;
  unsigned i;
  for (i = 1; i <= max_iterations + (max_iterations % 2) + 1; ++i) {
    counter = 0;
    ndrange(gx, gy, gz, tf, T, [&] { k_sdf::create_signed_distance_field(ping, pong, (int)i, &counter, (int)max_iterations); });
    std::swap(ping, pong);
    if (counter == 0 && i % 2 == 1) break;
  }
  memcpy(out, a.data(), N);
  return (int)i;
}

// app/reference_volume.cpp:22-41 around fetch_stats
void ref_fetch_stats(const int16_t* vol, int nx, int ny, int nz, int32_t stats_out[4], int threads) {
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  int stats[5] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MIN};
  ndrange(evenness(nx, 8), evenness(ny, 8), evenness(nz, 8), emu_tf{nullptr, 0}, nthreads(threads),
          [&] { k_figures::fetch_stats(&v, stats); });
  memcpy(stats_out, stats, 4 * sizeof(int));
}

// app/renderer.cpp:49-61 around tf_sort_values.  `bins` must have room for width*height + height + 2 entries: the kernel
// indexes x*height+y with x up to width; entries beyond width*height are the reference's out-of-bounds
// writes and are ignored by the caller.  Returns -1 if a voxel would index below 0 (not representable here).
int ref_histogram(const int16_t* vol, int nx, int ny, int nz, int width, int height, float min_v, float max_v, float min_g,
                  float max_g, uint32_t* bins, int threads) {
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  const size_t N = (size_t)nx * ny * nz;
  // guard only: a value below min_v makes x negative (the gradient axis cannot go below -height/2 + ... for the ranges
  // used here, and a negative y with x >= 1 stays inside the buffer; x == 0 with y < 0 is caught by min_g <= stats min)
  for (size_t i = 0; i < N; ++i)
    if ((float)vol[i] < min_v) return -1;
  {
    int st[4];
    ref_fetch_stats(vol, nx, ny, nz, st, threads);
    if ((float)st[2] < min_g) return -1;
  }
  memset(bins, 0, sizeof(uint32_t) * ((size_t)width * height + height + 2));
  ndrange(evenness(nx, 8), evenness(ny, 8), evenness(nz, 8), emu_tf{nullptr, 0}, nthreads(threads),
          [&] { k_histogram::tf_sort_values(&v, bins, width, height, min_v, max_v, min_g, max_g); });
  return 0;
}

// app/renderer.cpp:65-96 (host rounding pass + distinct sorted values) around tf_flush_color_frame
int ref_tf_color_frame(uint32_t* bins, int width, int height, uint8_t* out_rgba) {
  std::set<int, std::less<int>> possible_histories;
  for (unsigned y = 0; y < (unsigned)height; ++y)
    for (unsigned x = 0; x < (unsigned)width; ++x) {
      int value = bins[x * height + y];
      if (value != 0) {
        int roundingpart = std::max((int)(pow(10, (std::floor(std::log10(value))) - 1)), (int)1);
        int corrected_value = floor(value / roundingpart) * roundingpart;
        bins[x * height + y] = corrected_value;
        possible_histories.insert(corrected_value);
      }
    }
  memset(out_rgba, 0, (size_t)width * height * 4);
  if (possible_histories.size() == 0) return 0;
  std::vector<int> history(possible_histories.begin(), possible_histories.end());
  emu_image frame{out_rgba, width, height, 1, EMU_U8x4};
  ndrange(evenness(width, 16), evenness(height, 16), 1, emu_tf{nullptr, 0}, 1, [&] {
    k_histogram::tf_flush_color_frame(&frame, reinterpret_cast<int*>(bins), history.data(), (int)history.size());
  });
  return (int)history.size();
}

// app/reference_volume.cpp:70-80 around bilateral_filter
void ref_bilateral(const int16_t* vol, int nx, int ny, int nz, int16_t* out, int threads) {
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  memset(out, 0, (size_t)nx * ny * nz * 2);
  emu_image o{out, nx, ny, nz, EMU_S16};
  ndrange(evenness(nx, 8), evenness(ny, 8), evenness(nz, 8), emu_tf{nullptr, 0}, nthreads(threads),
          [&] { k_volume_filter::bilateral_filter(&v, &o); });
}

// app/reference_volume.cpp:54-68 around apply_clip
void ref_clip(const int16_t* vol, int nx, int ny, int nz, const int start[3], const int size[3], int16_t* out) {
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  emu_image o{out, size[0], size[1], size[2], EMU_S16};
  unsigned st[3] = {(unsigned)start[0], (unsigned)start[1], (unsigned)start[2]};
  unsigned len[4] = {(unsigned)size[0], (unsigned)size[1], (unsigned)size[2], 4};
  ndrange(evenness(size[0], 4), evenness(size[1], 4), evenness(size[2], 4), emu_tf{nullptr, 0}, 1,
          [&] { k_clip::apply_clip(&v, &o, st, len); });
}

// 2d_image_filter.cl:6-43 — no host call site exists in the reference (the kernel is dead code, "not tested"); launched here
// over exactly w x h work-items.  The kernel filters `frame` in place through a __read_write image, which makes a work-item's
// taps race with its neighbours' writes; reads are served from the input copy and writes go to `out` (cl_emu.hpp, emu_image::wdata).
void ref_image_filter2d(const uint8_t* rgba_in, int w, int h, int kernel_size, float sigma, uint8_t* rgba_out, int threads) {
  memcpy(rgba_out, rgba_in, (size_t)w * h * 4);
  emu_image f{const_cast<uint8_t*>(rgba_in), w, h, 1, EMU_U8x4, rgba_out};
  ndrange(w, h, 1, emu_tf{nullptr, 0}, nthreads(threads), [&] { k_image_filter::bilateral_filter(&f, kernel_size, sigma); });
}

// app/renderer.cpp:32-35 around buffer_reset
void ref_buffer_reset(uint16_t* cache, int nx, int ny, int nz) {
  emu_image v{nullptr, nx, ny, nz, EMU_S16};
  ndrange(evenness(nx, 4), evenness(ny, 4), evenness(nz, 4), emu_tf{nullptr, 0}, 1, [&] { k_buffer_reset::buffer_reset(&v, cache); });
}

// app/renderer.cpp:145-148 around `render`.  Work-items of the window [x0,x1) x [y0,y1) run in row-major order; with
// threads == 1 this is a deterministic single-phase execution (each pixel resolves right after its own add), which
// oracle.cpp reproduces with immediate=1.  frame must be W*H*4.
void ref_render_frame(const int16_t* vol, int nx, int ny, int nz, const int8_t* sdf, const uint8_t* env_rgba, int env_w,
                      int env_h, const emu_tf_rect* rects, int n_rects, uint16_t* cache, int W, int H, int x0, int y0, int x1,
                      int y1, const float cam_pos[3], const float cam_dir[3], int32_t random_seed, uint8_t* frame,
                      int threads) {
  emu_image v{const_cast<int16_t*>(vol), nx, ny, nz, EMU_S16};
  emu_image s{const_cast<int8_t*>(sdf), nx, ny, nz, EMU_S8};
  emu_image e{const_cast<uint8_t*>(env_rgba), env_w, env_h, 1, EMU_U8x4};
  emu_image f{frame, W, H, 1, EMU_U8x4};
  emu_tf tf{rects, n_rects};
  const int T = nthreads(threads);
#pragma omp parallel for schedule(dynamic, 2) num_threads(T)
  for (int y = y0; y < y1; ++y) {
    emu_tf_cur = tf;
    for (int x = x0; x < x1; ++x) {
      emu_wi.gid[0] = x; emu_wi.gid[1] = y; emu_wi.gid[2] = 0;
      k_ray_marching::render(&f, &v, &s, &e, cache, cam_pos[0], cam_pos[1], cam_pos[2], cam_dir[0], cam_dir[1], cam_dir[2],
                             random_seed);
    }
  }
}

}  // extern "C"
