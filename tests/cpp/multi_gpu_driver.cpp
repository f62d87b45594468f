// multi_gpu_driver.cpp — a torchrun-free C++ host that uses more than one GPU through nothing but the C-ABI (include/vr.h):
// one thread per rank, each with its own vr_ctx on its own device, NCCL communicator created by vr_comm_init.
//   1. sharded ingest            vr_volume_upload_sharded (every rank uploads its planes, the rest arrives over NVLink)
//   2. z-slab SDF build          vr_renderer_set_sharded_build + vr_renderer_flush
//   3. spp split (config 3)      seeds dealt round-robin, token cap 256/N, vr_cache_allreduce
//   4. image-tile split (config 4)  vr_renderer_set_row_blocks + vr_frame_allgather
//   5. z-slab histogram + volume filter   vr_histogram_sharded, vr_volume_filter_sharded
// Inputs / outputs are raw files so that tests/test_multirank.py can compare every result with the CPU oracle.
//   usage: driver <vol.raw> nx ny nz <env.raw> ew eh W H <seeds.bin> nseeds nranks block_rows <outdir>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include "../../include/vr.h"

static std::vector<char> slurp(const char* p) {
  std::ifstream f(p, std::ios::binary);
  return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const std::string& p, const void* d, size_t n) {
  std::ofstream f(p, std::ios::binary);
  f.write(reinterpret_cast<const char*>(d), (std::streamsize)n);
}
#define CHECK(call)                                                                                   \
  do {                                                                                                \
    if ((call) != VR_OK) {                                                                            \
      fprintf(stderr, "[rank %d] %s failed: %s\n", rank, #call, vr_last_error());                     \
      failed = 1;                                                                                     \
      return;                                                                                         \
    }                                                                                                 \
  } while (0)

static int failed = 0;

int main(int argc, char** argv) {
  if (argc != 15) { fprintf(stderr, "usage: see the header of multi_gpu_driver.cpp\n"); return 2; }
  const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
  const int ew = atoi(argv[6]), eh = atoi(argv[7]), W = atoi(argv[8]), H = atoi(argv[9]);
  const int nseeds = atoi(argv[11]), nranks = atoi(argv[12]), block_rows = atoi(argv[13]);
  const std::string out = argv[14];
  const auto vraw = slurp(argv[1]), eraw = slurp(argv[5]), sraw = slurp(argv[10]);
  if (vraw.size() != (size_t)nx * ny * nz * 2 || eraw.size() != (size_t)ew * eh * 4 || sraw.size() != (size_t)nseeds * 4) {
    fprintf(stderr, "input sizes do not match\n");
    return 2;
  }
  const int16_t* vox = reinterpret_cast<const int16_t*>(vraw.data());
  const int32_t* seeds = reinterpret_cast<const int32_t*>(sraw.data());
  uint8_t id[VR_COMM_ID_BYTES];
  if (vr_comm_unique_id(id) != VR_OK) { fprintf(stderr, "vr_comm_unique_id: %s\n", vr_last_error()); return 1; }
  const vr_tf_rect tf = {500.f, 1200.f, 0.f, 4000.f, 0, {255, 255, 255, 255}};  // ui.cpp:195
  const float s = (float)nx / 256.0f;
  const float pos[3] = {-200.f * s, 200.f * s, -200.f * s};
  const std::vector<char> draw = slurp((out + "/dir.bin").c_str());  // camera direction evaluated by the test (Position3D)
  if (draw.size() != 12) { fprintf(stderr, "dir.bin missing\n"); return 2; }
  float dir[3];
  memcpy(dir, draw.data(), 12);

  auto body = [&](int rank) {
    vr_ctx* ctx = nullptr;
    CHECK(vr_ctx_create(rank, &ctx));
    CHECK(vr_comm_init(ctx, rank, nranks, id));
    int z0 = 0, z1 = 0;
    CHECK(vr_comm_slab(ctx, nz, rank, &z0, &z1));
    vr_volume* vol = nullptr;
    CHECK(vr_volume_upload_sharded(ctx, vox + (size_t)nx * ny * z0, nx, ny, nz, &vol));
    vr_envmap* env = nullptr;
    CHECK(vr_envmap_bind(ctx, reinterpret_cast<const uint8_t*>(eraw.data()), ew, eh, &env));
    vr_renderer* r = nullptr;
    CHECK(vr_renderer_create(ctx, W, H, &r));
    CHECK(vr_renderer_set_scene(r, vol, env));
    CHECK(vr_renderer_set_tf(r, &tf, 1));
    CHECK(vr_renderer_set_sharded_build(r, 1));
    CHECK(vr_renderer_flush(r));
    std::vector<uint8_t> frame((size_t)W * H * 4);
    // ---- spp split: rank takes seeds rank, rank + N, ...
    std::vector<int32_t> mine;
    for (int k = rank; k < nseeds; k += nranks) mine.push_back(seeds[k]);
    CHECK(vr_renderer_set_token_cap(r, 256 / nranks > 0 ? 256 / nranks : 1));
    if (!mine.empty()) CHECK(vr_render_frames(r, pos, dir, mine.data(), (int)mine.size(), nullptr));
    CHECK(vr_cache_allreduce(r, frame.data()));
    if (rank == 0) {
      dump(out + "/frame_spp.bin", frame.data(), frame.size());
      std::vector<uint16_t> cache((size_t)nx * ny * nz * 4);
      CHECK(vr_cache_download(r, cache.data()));
      dump(out + "/cache_spp.bin", cache.data(), cache.size() * 2);
      std::vector<int8_t> sdf((size_t)nx * ny * nz);
      CHECK(vr_sdf_download(vr_renderer_sdf(r), sdf.data()));
      dump(out + "/sdf.bin", sdf.data(), sdf.size());
      std::vector<int16_t> back((size_t)nx * ny * nz);
      CHECK(vr_volume_download(vol, back.data()));
      dump(out + "/volume_gathered.bin", back.data(), back.size() * 2);
    }
    // ---- image-tile split: all seeds on every rank, rows dealt out in blocks
    CHECK(vr_renderer_reset_cache(r));
    CHECK(vr_renderer_set_token_cap(r, 256));
    CHECK(vr_renderer_set_row_blocks(r, block_rows, rank, nranks));
    CHECK(vr_render_frames(r, pos, dir, seeds, nseeds, nullptr));
    CHECK(vr_frame_allgather(r, frame.data()));
    if (rank == nranks - 1) dump(out + "/frame_tiles.bin", frame.data(), frame.size());
    // ---- z-slab histogram and filter
    int32_t st[4];
    CHECK(vr_volume_stats(vol, st));
    const float range[4] = {(float)st[0], (float)st[1], (float)st[2], (float)st[3]};
    std::vector<uint32_t> bins(100 * 80);
    CHECK(vr_histogram_sharded(vol, 100, 80, range, bins.data()));
    CHECK(vr_volume_filter_sharded(vol));
    if (rank == 0) {
      dump(out + "/bins.bin", bins.data(), bins.size() * 4);
      std::vector<int16_t> filtered((size_t)nx * ny * nz);
      CHECK(vr_volume_download(vol, filtered.data()));
      dump(out + "/filtered.bin", filtered.data(), filtered.size() * 2);
      printf("stats %d %d %d %d\n", st[0], st[1], st[2], st[3]);
    }
    // a second, explicitly sharded SDF build (threshold TF of tests/sdf/sdf_test.cpp:22) on the filtered volume
    const vr_tf_rect thr = {800.f, 0.f, 0.f, 0.f, VR_TF_THRESHOLD, {0, 0, 0, 0}};
    vr_sdf* sdf2 = nullptr;
    CHECK(vr_sdf_build_sharded(ctx, vol, &thr, 1, &sdf2));
    if (rank == 0) {
      std::vector<int8_t> sd((size_t)nx * ny * nz);
      CHECK(vr_sdf_download(sdf2, sd.data()));
      dump(out + "/sdf_thr_filtered.bin", sd.data(), sd.size());
    }
    // asynchronous sharded ingest (copy stream + its own communicator) while the compute stream is busy with a collective
    vr_volume* vol_b = nullptr;
    CHECK(vr_volume_upload_sharded_async(ctx, vox + (size_t)nx * ny * z0, nx, ny, nz, &vol_b));
    CHECK(vr_histogram_sharded(vol, 100, 80, range, bins.data()));
    if (rank == 0) {
      std::vector<int16_t> back((size_t)nx * ny * nz);
      int32_t st2[4];
      CHECK(vr_volume_stats(vol_b, st2));
      CHECK(vr_volume_download(vol_b, back.data()));
      dump(out + "/volume_gathered_async.bin", back.data(), back.size() * 2);
      printf("stats_async %d %d %d %d\n", st2[0], st2[1], st2[2], st2[3]);
    } else {
      CHECK(vr_volume_wait(vol_b));
    }
    CHECK(vr_comm_barrier(ctx));
    vr_volume_destroy(vol_b);
    vr_sdf_destroy(sdf2);
    vr_renderer_destroy(r);
    vr_envmap_destroy(env);
    vr_volume_destroy(vol);
    vr_ctx_destroy(ctx);
  };
  std::vector<std::thread> th;
  for (int k = 0; k < nranks; ++k) th.emplace_back(body, k);
  for (auto& t : th) t.join();
  if (failed) return 1;
  printf("EVERYTHING FINE (%d ranks)\n", nranks);
  return 0;
}
