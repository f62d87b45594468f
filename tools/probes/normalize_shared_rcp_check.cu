#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vr_device.cuh"
__global__ void k(unsigned long long n, unsigned long long* bad, unsigned seed0) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t s = hash_u32((uint32_t)i * 2654435761u + seed0), t = hash_u32(s + 12345u), u = hash_u32(t ^ 0x9e3779b9u);
    // mixes: integer-like directions (as in the RNG), blends of normals, random magnitudes
    f3 a;
    switch (i & 3) {
      case 0: a = {(float)((int)(s % 4096) - 3071), (float)((int)(t % 4096) - 3071), (float)((int)(u % 4096) - 3071)}; break;
      case 1: a = {__uint_as_float((s & 0x807fffffu) | ((100u + (t % 56u)) << 23)), __uint_as_float((t & 0x807fffffu) | ((100u + (u % 56u)) << 23)),
                   __uint_as_float((u & 0x807fffffu) | ((100u + (s % 56u)) << 23))}; break;
      case 2: a = {(float)(int)(s >> 16) * 1e-3f - 30.f, (float)(int)(t >> 20), 0.0f}; break;
      default: a = {__uint_as_float((s & 0x807fffffu) | (127u << 23)) * 0.3f, __uint_as_float((t & 0x807fffffu) | (126u << 23)), __uint_as_float((u & 0x807fffffu) | (120u << 23))}; break;
    }
    f3 p = normalize3(a), q = normalize3_shared_rcp(a);
    if (__float_as_uint(p.x) != __float_as_uint(q.x) || __float_as_uint(p.y) != __float_as_uint(q.y) || __float_as_uint(p.z) != __float_as_uint(q.z))
      atomicAdd(bad, 1ull);
  }
}
int main() {
  unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0;
  const unsigned long long n = 1ull << 33;
  k<<<148 * 16, 256>>>(n, bad, 7u);
  cudaDeviceSynchronize();
  printf("tested %llu vectors, mismatches %llu (%s)\n", n, *bad, cudaGetErrorString(cudaGetLastError()));
  return *bad != 0;
}
