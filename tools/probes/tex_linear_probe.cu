// tex_linear_probe.cu — does the CUDA texture unit (tex3D<float> on an int16 array, cudaReadModeNormalizedFloat, linear filter, border
// addressing, unnormalised coordinates) return what NVIDIA's OpenCL returns for read_imagei + CLK_FILTER_LINEAR on a SIGNED_INT16
// image (tests/golden/opencl_linear_probe.npz)?  Decides whether an opt-in "sample like the reference does on NVIDIA
// hardware" mode can be built on texture objects.
//   tex_linear_probe vol.i16 nx ny nz coords.f32 n out.f32
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k(cudaTextureObject_t t, const float* c, float* o, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = tex3D<float>(t, c[3 * i], c[3 * i + 1], c[3 * i + 2]);
}
int main(int argc, char** argv) {
  if (argc != 8) return 2;
  const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]), n = atoi(argv[6]);
  std::vector<short> vol((size_t)nx * ny * nz);
  std::vector<float> c((size_t)n * 3), o(n);
  FILE* f = fopen(argv[1], "rb"); if (!f || fread(vol.data(), 2, vol.size(), f) != vol.size()) return 3; fclose(f);
  f = fopen(argv[5], "rb"); if (!f || fread(c.data(), 4, c.size(), f) != c.size()) return 3; fclose(f);
  cudaArray_t arr;
  cudaChannelFormatDesc d = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindSigned);
  CK(cudaMalloc3DArray(&arr, &d, make_cudaExtent(nx, ny, nz)));
  cudaMemcpy3DParms p = {};
  p.srcPtr = make_cudaPitchedPtr(vol.data(), nx * 2, nx, ny);
  p.dstArray = arr; p.extent = make_cudaExtent(nx, ny, nz); p.kind = cudaMemcpyHostToDevice;
  CK(cudaMemcpy3D(&p));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModeLinear; td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float *dc, *dout;
  CK(cudaMalloc(&dc, c.size() * 4)); CK(cudaMalloc(&dout, (size_t)n * 4));
  CK(cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
  k<<<(n + 255) / 256, 256>>>(tex, dc, dout, n);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(o.data(), dout, (size_t)n * 4, cudaMemcpyDeviceToHost));
  f = fopen(argv[7], "wb"); fwrite(o.data(), 4, o.size(), f); fclose(f);
  return 0;
}
