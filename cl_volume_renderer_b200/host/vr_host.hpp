// vr_host.hpp — C++17 host side above the C-ABI (include/vr.h): the reference's scene and renderer classes with
// their names, constructor signatures, argument meaning and fail-hard error behaviour, so that app code written
// against cl-volume-renderer's `frame_emitter` / `renderer` keeps compiling and running with libvr.so underneath.
//
//   reference class (file:line)                                   here
//   clw_context            opencl_wrapper/include/clw_context.hpp:5-28   clw_context  (owns a vr_ctx)
//   volume_block           app/volume_block.hpp:6-34                     volume_block (unchanged data carrier)
//   image                  app/image.hpp:5-16                            image
//   Volume_Stats           app/reference_volume.hpp:8-23                 Volume_Stats
//   reference_volume       app/reference_volume.hpp:25-63, .cpp:11-112   reference_volume (owns a vr_volume)
//   signed_distance_field  app/signed_distance_field.hpp:5-12, .cpp:7-44 signed_distance_field (owns a vr_sdf)
//   env_map                app/env_map.hpp:6-15                          env_map (owns a vr_envmap)
//   Position3D, evenness   app/common.hpp:5-66                           same
//   ui_state, frame_emitter app/ui.hpp:14-37                             same (ui_state without the SDL members)
//   renderer               app/renderer.hpp:10-29, .cpp:8-158            renderer (owns a vr_renderer)
//   tf_rect_selection::create_cl_condition  app/tf_part.cpp:55-79        tf_rect_selection (no ImGui part)
//   ui::flush_tf           app/ui.cpp:160-168                            flush_tf()
//
// Errors: the reference prints the failing call and exit(1)s (clw_helper.hpp:293-309); so does vr_fail_hard().
#pragma once
#include <math.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/vr.h"

#define vr_fail_hard(call)                                                                             \
  do {                                                                                                 \
    int vr_status__ = (call);                                                                          \
    if (vr_status__ != VR_OK) {                                                                        \
      std::cerr << "Error, at: " << __FILE__ << ": " << __func__ << ": " << __LINE__ << '\n'           \
                << #call << " -> " << vr_status__ << ": " << vr_last_error() << '\n';                  \
      exit(1);                                                                                         \
    }                                                                                                  \
  } while (0)

#define SCREEN_WIDTH 2048  // app/common_defines.hpp:3-4
#define SCREEN_HEIGHT 1024

// ---- app/common.hpp:5-66 ------------------------------------------------------------------------------------------
struct Position3D {
  float val[3];
  Position3D(double alpha, double beta, double gamma, Position3D base) {
    val[0] = (cos(alpha) * cos(beta)) * base.val[0] + (cos(alpha) * sin(beta) - sin(alpha) * cos(gamma)) * base.val[1] +
             (cos(alpha) * sin(beta) * cos(gamma) + sin(alpha) * sin(gamma)) * base.val[2];
    val[1] = (-sin(beta)) * base.val[0] + (cos(beta) * sin(gamma)) * base.val[1] + (cos(beta) * cos(gamma)) * base.val[2];
    val[2] = (sin(alpha) * cos(beta)) * base.val[0] + (sin(alpha) * sin(beta) * sin(gamma) + cos(alpha) * cos(gamma)) * base.val[1] +
             (sin(alpha) * sin(beta) * cos(gamma) - cos(alpha) * sin(gamma)) * base.val[2];
    normalize();
  }
  Position3D(double x, double y, double z) { val[0] = x; val[1] = y; val[2] = z; }
  Position3D operator/(float other) { return {val[0] / other, val[1] / other, val[2] / other}; }
  double length() { return sqrtf(pow(val[0], 2) + pow(val[1], 2) + pow(val[2], 2)); }
  void normalize() { double len = length(); *this = *this / len; }
};

inline unsigned int evenness(const unsigned int g, const unsigned int l) {
  unsigned int mod = g % l;
  if (mod == 0) return g;
  return g + l - mod;
}

// ---- opencl_wrapper/include/clw_context.hpp:5-28 ----------------------------------------------------------------------
class clw_context {
 public:
  clw_context() { vr_fail_hard(vr_ctx_create(0, &m_ctx)); }  // device 0 (clw_context.cpp:40-47 picks platform 0 / device 0)
  explicit clw_context(int device_ordinal) { vr_fail_hard(vr_ctx_create(device_ordinal, &m_ctx)); }
  ~clw_context() { vr_ctx_destroy(m_ctx); }
  clw_context(const clw_context&) = delete;
  clw_context& operator=(const clw_context&) = delete;
  vr_ctx* get() const { return m_ctx; }

 private:
  vr_ctx* m_ctx = nullptr;
};

// ---- app/volume_block.hpp:6-34, app/image.hpp:5-16 ---------------------------------------------------------------------
struct volume_block {
  std::vector<short> m_voxels;
  const unsigned int m_voxel_count_x, m_voxel_count_y, m_voxel_count_z;
  const float m_voxel_size_x, m_voxel_size_y, m_voxel_size_z;
  volume_block(const unsigned int cx, const unsigned int cy, const unsigned int cz, const float sx, const float sy, const float sz)
      : m_voxel_count_x(cx), m_voxel_count_y(cy), m_voxel_count_z(cz), m_voxel_size_x(sx), m_voxel_size_y(sy), m_voxel_size_z(sz) {}
  volume_block(std::vector<short>&& v, const unsigned int cx, const unsigned int cy, const unsigned int cz, const float sx,
               const float sy, const float sz)
      : m_voxels(v), m_voxel_count_x(cx), m_voxel_count_y(cy), m_voxel_count_z(cz), m_voxel_size_x(sx), m_voxel_size_y(sy),
        m_voxel_size_z(sz) {}
};

struct image {
  std::vector<unsigned char> m_pixels;
  const unsigned int m_width, m_height, m_pixel_depth;
  image(std::vector<unsigned char>&& v, const unsigned int w, const unsigned int h, const unsigned int d)
      : m_pixels(v), m_width(w), m_height(h), m_pixel_depth(d) {}
};

// ---- app/reference_volume.hpp:8-63 --------------------------------------------------------------------------------------
struct Volume_Stats {
  float min_v, max_v, min_g, max_g;
  template <typename T>
  Volume_Stats(const T& value_stats, const T& gradient_stats) {
    min_v = value_stats[0]; max_v = value_stats[1]; min_g = gradient_stats[0]; max_g = gradient_stats[1];
  }
  Volume_Stats() { min_v = max_v = min_g = max_g = 0; }
};

class reference_volume {
 public:
  // reference_volume.cpp:11-44: uploads the voxels and fetches min/max value and gradient
  reference_volume(clw_context& c, volume_block* b)
      : ctx(c), volume_size({b->m_voxel_count_x, b->m_voxel_count_y, b->m_voxel_count_z}), cropped_volume_size(volume_size) {
    if ((size_t)b->m_voxel_count_x * b->m_voxel_count_y * b->m_voxel_count_z != b->m_voxels.size()) {
      std::cerr << "Error, moved array is not of the correct size\n";  // clw_image.hpp:37-42
      exit(1);
    }
    vr_fail_hard(vr_volume_upload(ctx.get(), b->m_voxels.data(), (int)volume_size[0], (int)volume_size[1], (int)volume_size[2], &vol));
    int32_t s[4];
    vr_fail_hard(vr_volume_stats(vol, s));
    value_range = {s[0], s[1]};
    gradient_range = {s[2], s[3]};
    std::cout << gradient_range[1];  // reference_volume.cpp:42
  }
  ~reference_volume() { vr_volume_destroy(vol); }
  reference_volume(const reference_volume&) = delete;
  reference_volume& operator=(const reference_volume&) = delete;

  void set_value_clip(std::array<int, 2> clip) { value_clip = clip; vr_fail_hard(vr_volume_set_value_clip(vol, clip[0], clip[1])); }
  void set_gradient_clip(std::array<int, 2> clip) { gradient_clip = clip; vr_fail_hard(vr_volume_set_gradient_clip(vol, clip[0], clip[1])); }
  void set_clipping(std::array<size_t, 3> min, std::array<size_t, 3> max) {  // reference_volume.cpp:54-68
    assert(min[0] < max[0]); assert(min[1] < max[1]); assert(min[2] < max[2]);
    const uint32_t mn[3] = {(uint32_t)min[0], (uint32_t)min[1], (uint32_t)min[2]};
    const uint32_t mx[3] = {(uint32_t)max[0], (uint32_t)max[1], (uint32_t)max[2]};
    vr_fail_hard(vr_volume_clip(vol, mn, mx));
    cropped_volume_size = {max[0] - min[0], max[1] - min[1], max[2] - min[2]};
  }
  void filter() { vr_fail_hard(vr_volume_filter(vol)); }  // reference_volume.cpp:70-80
  // VR_SAMPLING_NEAREST (default) or VR_SAMPLING_HW_LINEAR: how fetch_stats (recomputed now), the TF histogram and filter() read the
  // volume — the filter OpenCL defines for integer images, or what NVIDIA hardware does with the reference's samplers (include/vr.h)
  void set_sampling(int mode) { vr_fail_hard(vr_volume_set_sampling(vol, mode)); }
  std::array<int, 2> get_value_range() const { return {std::max(value_clip[0], value_range[0]), std::min(value_clip[1], value_range[1])}; }
  std::array<int, 2> get_gradient_range() const {
    return {std::max(gradient_clip[0], gradient_range[0]), std::min(gradient_clip[1], gradient_range[1])};
  }
  const std::array<size_t, 3>& get_original_volume_size() const { return volume_size; }
  const std::array<size_t, 3>& get_volume_size() const { return cropped_volume_size; }
  std::array<size_t, 3> get_volume_size_evenness(unsigned int l) const {
    return {evenness(cropped_volume_size[0], l), evenness(cropped_volume_size[1], l), evenness(cropped_volume_size[2], l)};
  }
  size_t get_volume_length() const { return cropped_volume_size[0] * cropped_volume_size[1] * cropped_volume_size[2]; }
  Volume_Stats get_volume_stats() const { return Volume_Stats(get_value_range(), get_gradient_range()); }
  // replaces get_reference_volume(): the device image is now an opaque handle
  vr_volume* get_reference_volume() const { return vol; }
  clw_context& context() const { return ctx; }

 private:
  clw_context& ctx;
  std::array<size_t, 3> volume_size, cropped_volume_size;
  vr_volume* vol = nullptr;
  std::array<int, 2> value_range{0, 0}, gradient_range{0, 0};
  std::array<int, 2> value_clip = {std::numeric_limits<int>::min(), std::numeric_limits<int>::max()};
  std::array<int, 2> gradient_clip = {std::numeric_limits<int>::min(), std::numeric_limits<int>::max()};
};

// ---- app/signed_distance_field.hpp:5-12 -----------------------------------------------------------------------------------
class signed_distance_field {
 public:
  explicit signed_distance_field(clw_context&) {}  // dummy 2x2x2 field in the reference (signed_distance_field.cpp:42-45)
  // signed_distance_field.cpp:7-35; local_cl_code is the generated `is_event_gen` source
  signed_distance_field(clw_context& c, const reference_volume& rv, std::string local_cl_code) : size(rv.get_volume_size()) {
    vr_tf_rect rects[VR_TF_MAX_RECTS];
    int n = 0;
    vr_fail_hard(vr_tf_parse(local_cl_code.c_str(), rects, VR_TF_MAX_RECTS, &n));
    vr_fail_hard(vr_sdf_build(c.get(), rv.get_reference_volume(), rects, n, &sdf));
  }
  ~signed_distance_field() { vr_sdf_destroy(sdf); }
  signed_distance_field(const signed_distance_field&) = delete;
  signed_distance_field& operator=(const signed_distance_field&) = delete;
  // clw_image<char>::pull() + operator[] (tests/sdf/sdf_test.cpp:24-31): host copy, x fastest
  std::vector<char> pull() const {
    std::vector<char> out(size[0] * size[1] * size[2]);
    if (sdf) vr_fail_hard(vr_sdf_download(sdf, reinterpret_cast<int8_t*>(out.data())));
    return out;
  }
  vr_sdf* get_sdf_buffer() { return sdf; }

 private:
  vr_sdf* sdf = nullptr;
  std::array<size_t, 3> size{0, 0, 0};
};

// ---- app/env_map.hpp:6-15 -----------------------------------------------------------------------------------------------
class env_map {
 public:
  env_map(clw_context& ctx, image& env) {
    if ((size_t)env.m_width * env.m_height * 4 != env.m_pixels.size()) {
      std::cerr << "Error, moved array is not of the correct size\n";
      exit(1);
    }
    vr_fail_hard(vr_envmap_bind(ctx.get(), env.m_pixels.data(), (int)env.m_width, (int)env.m_height, &map));
  }
  ~env_map() { vr_envmap_destroy(map); }
  env_map(const env_map&) = delete;
  env_map& operator=(const env_map&) = delete;
  vr_envmap* get_buffer() const { return map; }

 private:
  vr_envmap* map = nullptr;
};

// ---- app/ui.hpp:14-37 ---------------------------------------------------------------------------------------------------
struct ui_state {
  std::string path;
  bool path_changed;
  int height;
  int width;
  Position3D position;
  float direction_look[2];
  bool cam_changed;
};

class frame_emitter {
 public:
  virtual ~frame_emitter() {}
  virtual void image_set(const reference_volume* volume, const env_map* map) = 0;
  virtual void next_event_code_set(const std::string cl_code) = 0;
  virtual void flush_changes() = 0;
  virtual void* render_frame(struct ui_state& state, bool& frame_changed) = 0;
  virtual void* render_tf(const unsigned int width, const unsigned int height) = 0;
};

// ---- app/tf_part.cpp:8-16,55-79 and app/ui.cpp:160-168 ---------------------------------------------------------------------
class tf_selection {
 public:
  virtual ~tf_selection() {}
  virtual std::string create_cl_condition(Volume_Stats stats) = 0;
};

class tf_rect_selection : public tf_selection {
 public:
  float min_v, max_v, min_g, max_g;
  float color[4];
  tf_rect_selection(unsigned int id, float min_v, float max_v, float min_g, float max_g)
      : min_v(min_v), max_v(max_v), min_g(min_g), max_g(max_g), id(id * 100) {
    color[0] = color[1] = color[2] = color[3] = 1.0;
  }
  std::string create_cl_condition(Volume_Stats stats) override {
    std::ostringstream cl_code;
    cl_code << "  if(value >= " << min_v << " && value <= " << max_v;
    if (min_g > stats.min_g || max_g < stats.max_g) cl_code << " && gradient > " << min_g << " && gradient < " << max_g;
    cl_code << ")\n {\n";
    cl_code << "    int4 tmp_color = {" << (int)(color[0] * 255) << "," << (int)(color[1] * 255) << "," << (int)(color[2] * 255) << ","
            << (int)(color[3] * 255) << "};\n";
    cl_code << "    *color = tmp_color;\n    return true;\n }\n";
    return cl_code.str();
  }

 private:
  unsigned int id;
};

inline void flush_tf(frame_emitter* emitter, Volume_Stats stats, std::vector<tf_selection*> selection) {
  std::string cl_code = "inline bool is_event_gen(short value, short gradient, int4 *color){\n";
  for (auto s : selection) cl_code += s->create_cl_condition(stats);
  cl_code += "  \n  return false;\n}\n";
  emitter->next_event_code_set(cl_code);
}

// ---- app/renderer.hpp:10-29, app/renderer.cpp:8-158 ---------------------------------------------------------------------------
class renderer : public frame_emitter {
 public:
  // renderer.cpp:8-17: the reference allocates a SCREEN_WIDTH x SCREEN_HEIGHT frame image; the frame actually rendered is
  // state.width x state.height, so the device renderer is (re)created for that size on first use.
  explicit renderer(clw_context& c) : ctx(c) {}
  ~renderer() override { vr_renderer_destroy(r); }
  void image_set(const reference_volume* rv, const env_map* map) override { volume = rv; emap = map; }
  void next_event_code_set(const std::string cl_code) override { local_cl_code = cl_code; }
  void flush_changes() override {  // renderer.cpp:25-43
    flush_pending = true;
    if (r) do_flush();
  }
  void* render_frame(struct ui_state& state, bool& frame_changed) override {  // renderer.cpp:131-158
    frame_changed = false;
    ensure(state.width, state.height);
    if (!state.cam_changed && !state.path_changed) return vr_renderer_host_frame(r);
    Position3D vec(state.direction_look[0], state.direction_look[1], 0.0, {1.0, 0.0, 0.0});
    int random_seed = std::rand();
    const float pos[3] = {state.position.val[0], state.position.val[1], state.position.val[2]};
    vr_fail_hard(vr_render_frame(r, pos, vec.val, random_seed, vr_renderer_host_frame(r)));
    state.cam_changed = false;
    state.path_changed = false;
    frame_changed = true;
    return vr_renderer_host_frame(r);
  }
  // renderer.cpp:45 names its parameters (height, width); both call sites pass 500,500 (ui.cpp:148,153)
  void* render_tf(const unsigned int height, const unsigned int width) override {
    ensure(rw ? rw : SCREEN_WIDTH, rh ? rh : SCREEN_HEIGHT);
    tfframe.resize((size_t)width * height * 4);
    vr_fail_hard(vr_render_tf(r, (int)width, (int)height, tfframe.data()));
    return tfframe.data();
  }
  // opencl_kernels/2d_image_filter.cl:6-43 `bilateral_filter(frame, kernel_size, sigma)`: the reference ships the kernel but has
  // no host method that launches it; this is the one it would take.  Filters the frame last rendered and returns the same
  // pointer render_frame() returns.  mode: VR_FILTER2D_REFERENCE (the kernel as written) or VR_FILTER2D_BILATERAL.
  void* filter_frame(int kernel_size, float sigma, int mode = VR_FILTER2D_REFERENCE) {
    ensure(rw ? rw : SCREEN_WIDTH, rh ? rh : SCREEN_HEIGHT);
    vr_fail_hard(vr_renderer_filter_frame(r, kernel_size, sigma, mode, vr_renderer_host_frame(r)));
    return vr_renderer_host_frame(r);
  }
  // VR_SAMPLING_NEAREST (default: the filter OpenCL defines for the reference's integer images) or VR_SAMPLING_HW_LINEAR (what
  // NVIDIA hardware does with the reference's CLK_FILTER_LINEAR samplers); takes effect at the next flush_changes()
  void set_sampling(int mode) {
    sampling = mode;
    if (r) vr_fail_hard(vr_renderer_set_sampling(r, mode));
  }
  vr_renderer* handle() const { return r; }

 private:
  void ensure(int w, int h) {
    if (r && w == rw && h == rh) return;
    if (r) vr_renderer_destroy(r);
    vr_fail_hard(vr_renderer_create(ctx.get(), w, h, &r));
    vr_fail_hard(vr_renderer_set_sampling(r, sampling));
    // the UI calls render_frame once per sample while the camera rests (ui.cpp:296): keep the seed-independent primary
    // segment between calls; camera moves, flushes and row-window changes re-march (include/vr.h)
    vr_fail_hard(vr_renderer_set_primary_reuse(r, 2));
    rw = w; rh = h;
    if (flush_pending || flushed_once) do_flush();
    else if (volume && emap) vr_fail_hard(vr_renderer_set_scene(r, volume->get_reference_volume(), emap->get_buffer()));
  }
  void do_flush() {
    vr_fail_hard(vr_renderer_set_scene(r, volume->get_reference_volume(), emap->get_buffer()));
    vr_fail_hard(vr_renderer_set_tf_code(r, local_cl_code.c_str()));
    vr_fail_hard(vr_renderer_flush(r));
    flush_pending = false;
    flushed_once = true;
  }
  clw_context& ctx;
  vr_renderer* r = nullptr;
  int rw = 0, rh = 0;
  int sampling = VR_SAMPLING_NEAREST;
  bool flush_pending = false, flushed_once = false;
  std::vector<unsigned char> tfframe;
  const reference_volume* volume = nullptr;
  const env_map* emap = nullptr;
  std::string local_cl_code;
};
