// vr_headless — headless replacement of app/main.cpp + app/ui.cpp: loads a NRRD volume and an
// environment map, sets up the scene exactly like ui::run (app/ui.cpp:170-199), accumulates --spp frames with the seeds
// std::rand() would give the reference (app/renderer.cpp:142) and writes the last frame as a binary PPM (upright: the
// frame's row 0 is the bottom of the view).
//
//   vr_headless <volume.nrrd> <envmap.hdr|.ppm|.WxH.rgba> [--w 1920] [--h 1080] [--spp 64] [--out frame.ppm]
//               [--pos x y z] [--look a b] [--tf "min_v,max_v,min_g,max_g,r,g,b,a;..."] [--filter] [--clip x0 y0 z0 x1 y1 z1]
//               [--tf-image tf.ppm] [--raw frame.rgba] [--frame-filter kernel_size sigma reference|bilateral]
//               [--sampling nearest|hw-linear]
// --sampling hw-linear: value / gradient / environment reads interpolated by the texture unit, as NVIDIA hardware executes the
// reference's CLK_FILTER_LINEAR samplers (default nearest: the filter OpenCL defines for integer images).
// --frame-filter runs opencl_kernels/2d_image_filter.cl over the final frame ("reference" = the kernel as written).
#include <cstring>
#include <fstream>

#include "vr_io.hpp"

static void write_ppm(const std::string& path, const unsigned char* rgba, int w, int h, bool flip) {
  std::ofstream f(path, std::ios::binary);
  f << "P6\n" << w << " " << h << "\n255\n";
  std::vector<unsigned char> row((size_t)w * 3);
  for (int y = 0; y < h; ++y) {
    const unsigned char* src = rgba + (size_t)(flip ? h - 1 - y : y) * w * 4;
    for (int x = 0; x < w; ++x) { row[3 * x] = src[4 * x]; row[3 * x + 1] = src[4 * x + 1]; row[3 * x + 2] = src[4 * x + 2]; }
    f.write(reinterpret_cast<const char*>(row.data()), row.size());
  }
}

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cout << "Error, required 2 parameters, nrrd & hdre.";  // main.cpp:10
    return -1;
  }
  int W = 1920, H = 1080, spp = 64;
  std::string out = "frame.ppm", raw_out, tf_image, tf_spec;
  double pos[3] = {0, 0, 0}, look[2] = {0.9, 6.183};  // ui.cpp:178
  bool have_pos = false, filter = false, clip = false;
  int ff_k = -1, ff_mode = VR_FILTER2D_REFERENCE, sampling = VR_SAMPLING_NEAREST;
  float ff_sigma = 1.0f;
  size_t cmin[3] = {0, 0, 0}, cmax[3] = {0, 0, 0};
  for (int i = 3; i < argc; ++i) {
    const std::string a = argv[i];
    auto need = [&](int n) { if (i + n >= argc) { std::cerr << "missing value for " << a << '\n'; exit(2); } };
    if (a == "--w") { need(1); W = atoi(argv[++i]); }
    else if (a == "--h") { need(1); H = atoi(argv[++i]); }
    else if (a == "--spp") { need(1); spp = atoi(argv[++i]); }
    else if (a == "--out") { need(1); out = argv[++i]; }
    else if (a == "--raw") { need(1); raw_out = argv[++i]; }
    else if (a == "--tf-image") { need(1); tf_image = argv[++i]; }
    else if (a == "--tf") { need(1); tf_spec = argv[++i]; }
    else if (a == "--filter") filter = true;
    else if (a == "--sampling") {
      need(1);
      const std::string m = argv[++i];
      if (m != "nearest" && m != "hw-linear") { std::cerr << "--sampling must be nearest or hw-linear\n"; return 2; }
      sampling = m == "nearest" ? VR_SAMPLING_NEAREST : VR_SAMPLING_HW_LINEAR;
    }
    else if (a == "--frame-filter") {
      need(3);
      ff_k = atoi(argv[++i]);
      ff_sigma = (float)atof(argv[++i]);
      const std::string m = argv[++i];
      if (m != "reference" && m != "bilateral") { std::cerr << "--frame-filter mode must be reference or bilateral\n"; return 2; }
      ff_mode = m == "reference" ? VR_FILTER2D_REFERENCE : VR_FILTER2D_BILATERAL;
    }
    else if (a == "--pos") { need(3); for (int k = 0; k < 3; ++k) pos[k] = atof(argv[++i]); have_pos = true; }
    else if (a == "--look") { need(2); look[0] = atof(argv[++i]); look[1] = atof(argv[++i]); }
    else if (a == "--clip") { need(6); for (int k = 0; k < 3; ++k) cmin[k] = atol(argv[++i]); for (int k = 0; k < 3; ++k) cmax[k] = atol(argv[++i]); clip = true; }
    else { std::cerr << "unknown option " << a << '\n'; return 2; }
  }

  clw_context ctx;
  renderer render_ctx(ctx);
  frame_emitter* emitter = &render_ctx;
  render_ctx.set_sampling(sampling);

  nrrd_loader vloader;
  volume_block v = vloader.load_file(argv[1]);
  const double scale = v.m_voxel_count_x / 256.0;
  reference_volume rv(ctx, &v);
  rv.set_sampling(sampling);
  rv.set_value_clip({-2000, 3000});
  rv.set_gradient_clip({0, 4000});
  if (clip) rv.set_clipping({cmin[0], cmin[1], cmin[2]}, {cmax[0], cmax[1], cmax[2]});
  if (filter) rv.filter();

  hdre_loader iloader;
  image em = iloader.load_file(argv[2]);
  env_map emap(ctx, em);
  emitter->image_set(&rv, &emap);

  std::vector<tf_selection*> selection;
  if (tf_spec.empty()) {
    selection.push_back(new tf_rect_selection(0, 500.f, 1200.f, 0.0f, 4000.f));  // ui.cpp:195
  } else {
    unsigned id = 0;
    for (const std::string& part : vr_io_detail::split(tf_spec, ';')) {
      const std::vector<std::string> t = vr_io_detail::split(part, ',');
      if (t.size() != 8) { std::cerr << "--tf wants min_v,max_v,min_g,max_g,r,g,b,a per rectangle\n"; return 2; }
      auto* r = new tf_rect_selection(id++, std::stof(t[0]), std::stof(t[1]), std::stof(t[2]), std::stof(t[3]));
      for (int k = 0; k < 4; ++k) r->color[k] = std::stof(t[4 + k]);
      selection.push_back(r);
    }
  }
  flush_tf(emitter, rv.get_volume_stats(), selection);
  emitter->flush_changes();

  if (!tf_image.empty()) write_ppm(tf_image, static_cast<unsigned char*>(emitter->render_tf(500, 500)), 500, 500, false);

  if (!have_pos) { pos[0] = -200 * scale; pos[1] = 200 * scale; pos[2] = -200 * scale; }
  struct ui_state state = {argv[1], true, H, W, {pos[0], pos[1], pos[2]}, {(float)look[0], (float)look[1]}, true};
  void* frame = nullptr;
  for (int k = 0; k < spp; ++k) {
    bool changed = false;
    state.cam_changed = true;
    frame = emitter->render_frame(state, changed);
  }
  if (frame && ff_k >= 0) frame = render_ctx.filter_frame(ff_k, ff_sigma, ff_mode);
  if (frame) {
    write_ppm(out, static_cast<unsigned char*>(frame), W, H, true);
    if (!raw_out.empty()) {
      std::ofstream f(raw_out, std::ios::binary);
      f.write(static_cast<const char*>(frame), (size_t)W * H * 4);
    }
  }
  const Volume_Stats st = rv.get_volume_stats();
  std::cout << "\nvolume " << rv.get_volume_size()[0] << "x" << rv.get_volume_size()[1] << "x" << rv.get_volume_size()[2] << " stats "
            << st.min_v << " " << st.max_v << " " << st.min_g << " " << st.max_g << " spp " << spp << " -> " << out << '\n';
  return 0;
}
