// frame_emitter_driver.cpp — drives the C++ host shim exactly the way the reference's app does, headless:
// ui::run start-up (app/ui.cpp:170-199), the per-frame call (ui.cpp:296) and the SDF test (tests/sdf/sdf_test.cpp:12-34).
// Inputs/outputs are raw files so that tests/test_host_shim_gpu.py can compare against the CPU oracle.
//   usage: driver <vol.raw> nx ny nz <env.raw> ew eh W H frames <outdir>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "../../cl_volume_renderer_b200/host/vr_host.hpp"

static std::vector<char> slurp(const char* p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const std::string& p, const void* d, size_t n) {
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(d), n);
}

int main(int argc, char** argv) {
  if (argc != 12) { std::cerr << "usage\n"; return 2; }
  const unsigned nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
  const unsigned ew = atoi(argv[6]), eh = atoi(argv[7]);
  const int W = atoi(argv[8]), H = atoi(argv[9]), frames = atoi(argv[10]);
  const std::string out = argv[11];

  clw_context ctx;                       // main.cpp:13
  renderer render_ctx(ctx);              // main.cpp:14
  frame_emitter* emitter = &render_ctx;  // ui::run(frame_emitter*)

  auto raw = slurp(argv[1]);
  std::vector<short> vox(raw.size() / 2);
  memcpy(vox.data(), raw.data(), vox.size() * 2);
  volume_block v(std::move(vox), nx, ny, nz, 1.f, 1.f, 1.f);
  reference_volume rv(ctx, &v);          // ui.cpp:186
  rv.set_value_clip({-2000, 3000});      // ui.cpp:187
  rv.set_gradient_clip({0, 4000});       // ui.cpp:188

  auto eraw = slurp(argv[5]);
  image em(std::vector<unsigned char>(eraw.begin(), eraw.end()), ew, eh, 4);
  env_map emap(ctx, em);                 // ui.cpp:192
  emitter->image_set(&rv, &emap);        // ui.cpp:193
  std::vector<tf_selection*> selection;
  selection.push_back(new tf_rect_selection(0, 500.f, 1200.f, 0.0f, 4000.f));  // ui.cpp:195
  flush_tf(emitter, rv.get_volume_stats(), selection);                         // ui.cpp:196
  emitter->flush_changes();                                                    // ui.cpp:197

  struct ui_state state = {argv[1], true, H, W, {-200.0 * nx / 256.0, 200.0 * nx / 256.0, -200.0 * nx / 256.0}, {0.9f, 6.183f}, true};
  void* frame = nullptr;
  for (int k = 0; k < frames; ++k) {
    bool changed = false;
    state.cam_changed = true;            // keep accumulating, like a key press would
    frame = emitter->render_frame(state, changed);  // ui.cpp:296
    if (!changed) { std::cerr << "frame_changed not set\n"; return 1; }
  }
  bool changed = true;
  void* same = emitter->render_frame(state, changed);  // no change requested: cached host frame, frame_changed=false
  if (changed || same != frame) { std::cerr << "early-out contract broken (renderer.cpp:134-135)\n"; return 1; }
  dump(out + "/frame.bin", frame, (size_t)W * H * 4);

  void* tf = emitter->render_tf(100, 80);  // ui.cpp:148 uses 500x500; (height,width) as in renderer.cpp:45
  dump(out + "/tf.bin", tf, (size_t)100 * 80 * 4);

  // tests/sdf/sdf_test.cpp:22-31
  signed_distance_field sdf(ctx, rv, "inline bool is_event_gen(short value, short gradient, uint4 *color){ return (value > 800); }");
  auto s = sdf.pull();
  dump(out + "/sdf.bin", s.data(), s.size());
  auto st = rv.get_volume_stats();
  std::cout << "\nstats " << st.min_v << " " << st.max_v << " " << st.min_g << " " << st.max_g << "\nEVERYTHING FINE\n";
  return 0;
}
