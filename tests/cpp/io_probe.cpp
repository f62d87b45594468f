// io_probe — CPU-only check of the ingest code (vr_io.hpp): prints dims, voxel sizes and checksums of a NRRD file and an
// environment map so that tests/test_io_cpu.py can compare them with numpy.
#include "../../cl_volume_renderer_b200/host/vr_io.hpp"
int main(int argc, char** argv) {
  if (argc < 3) return 2;
  if (std::string(argv[1]) == "nrrd") {
    nrrd_loader l;
    volume_block b = l.load_file(argv[2]);
    long long sum = 0, wsum = 0;
    for (size_t i = 0; i < b.m_voxels.size(); ++i) { sum += b.m_voxels[i]; wsum += (long long)b.m_voxels[i] * (long long)(i % 1009); }
    printf("%u %u %u %g %g %g %lld %lld %d\n", b.m_voxel_count_x, b.m_voxel_count_y, b.m_voxel_count_z, b.m_voxel_size_x,
           b.m_voxel_size_y, b.m_voxel_size_z, sum, wsum, (int)b.m_voxels[0]);
  } else {
    hdre_loader l;
    image im = l.load_file(argv[2]);
    if (argc > 3) { FILE* f = fopen(argv[3], "wb"); fwrite(im.m_pixels.data(), 1, im.m_pixels.size(), f); fclose(f); }
    printf("%u %u %u\n", im.m_width, im.m_height, im.m_pixel_depth);
  }
  return 0;
}
