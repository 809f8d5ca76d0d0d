// Is the texture unit's unorm8 -> float conversion exactly RN(c / 255)?  (would let the colour filter square the fetched
// float instead of looking (c/255)^2 up in shared memory).  build: nvcc -gencode arch=compute_100a,code=sm_100a -o unorm_exact unorm_exact.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
__global__ void k(cudaTextureObject_t t, float *out) {
  const int c = threadIdx.x;
  float4 v = tex2D<float4>(t, (c + 0.5f) / 256.0f, 0.5f);
  out[c] = v.x;
  float4 g;
  asm volatile("tld4.r.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w)
               : "l"(t), "f"((c + 1.0f) / 256.0f), "f"(1.0f / 2.0f));
  out[256 + c] = g.w;  // (i0, j0) with the point at the common corner of texels c, c+1
}
int main() {
  unsigned char h[2][256][4];
  for (int y = 0; y < 2; ++y)
    for (int c = 0; c < 256; ++c) h[y][c][0] = h[y][c][1] = h[y][c][2] = h[y][c][3] = (unsigned char)c;
  cudaChannelFormatDesc d = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
  cudaArray_t a;
  cudaMallocArray(&a, &d, 256, 2, cudaArrayTextureGather);
  cudaMemcpy2DToArray(a, 0, 0, h, 1024, 1024, 2, cudaMemcpyHostToDevice);
  cudaResourceDesc rd; memset(&rd, 0, sizeof rd); rd.resType = cudaResourceTypeArray; rd.res.array.array = a;
  cudaTextureDesc td; memset(&td, 0, sizeof td);
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat; td.normalizedCoords = 1;
  cudaTextureObject_t t; cudaCreateTextureObject(&t, &rd, &td, NULL);
  float *out, res[512]; cudaMalloc(&out, 2048);
  k<<<1, 256>>>(t, out);
  cudaMemcpy(res, out, 2048, cudaMemcpyDeviceToHost);
  int bad = 0, bad4 = 0;
  for (int c = 0; c < 256; ++c) {
    volatile float want = (float)c / 255.0f;
    if (res[c] != want) { if (bad < 5) printf("tex2D c=%d got %.9g want %.9g\n", c, res[c], want); ++bad; }
    if (res[256 + c] != want) { if (bad4 < 5) printf("tld4 c=%d got %.9g want %.9g\n", c, res[256 + c], want); ++bad4; }
  }
  printf("unorm8->float: tex2D mismatches %d / 256, tld4 mismatches %d / 256 (%s)\n", bad, bad4, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
