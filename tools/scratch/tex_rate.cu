// Microbenchmark (development aid, not part of the library): issue rate of the instructions the column-parallel march
// is made of, on one B200 -- tld4 on an R16F texture (L1-resident footprint; march-like line footprint through L2), tex.2d
// point fetches, FRND, and the plain FP32 pipe for scale.  Prints warp instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/tex_rate tex_rate.cu && bin/tex_rate
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_tld4(unsigned long long tex, int iters, float step, float inv, float *sink, int mode) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  // mode 0: every warp walks the same 32x32 texel patch (L1 hits); mode 1: a line across the map, 2 texels per lane,
  // advancing per iteration (march at far depth); mode 2: 0.1 texel per lane (near depth)
  float x = mode == 0 ? (float)(lane & 7) : (float)(tid % 4096) * (mode == 1 ? 2.0f : 0.1f);
  float y = mode == 0 ? (float)(lane >> 3) : (float)((tid / 4096) * 7 % 2048);
  float acc = 0.f;
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a, b, c, d;
      const float xx = x + (mode == 0 ? (float)u : step * (float)(i + u)), yy = y + (mode == 0 ? 0.f : 0.37f * (float)(i + u));
      asm volatile("tld4.r.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(tex), "f"(xx * inv), "f"(yy * inv));
      acc += a + b + c + d;
    }
  }
  if (acc == 123.456f) *sink = acc;
}
__global__ void k_texpoint(unsigned long long tex, int iters, float inv, float *sink) {
  const int lane = threadIdx.x & 31;
  float x = (float)(lane & 7), y = (float)(lane >> 3), acc = 0.f;
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a, b, c, d;
      asm volatile("tex.2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6}];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(tex), "f"((x + u + 0.5f) * inv), "f"((y + 0.5f) * inv));
      acc += a;
    }
  }
  if (acc == 123.456f) *sink = acc;
}
__global__ void k_frnd(int iters, float seed, float *sink) {
  float a = seed + threadIdx.x, b = a + 0.3f, c = a + 0.7f, d = a + 1.1f;
  for (int i = 0; i < iters; i += 4) {
    a = floorf(a * 1.0001f); b = floorf(b * 1.0001f); c = floorf(c * 1.0001f); d = floorf(d * 1.0001f);
  }
  if (a + b + c + d == 123.456f) *sink = a;
}
__global__ void k_fmul(int iters, float seed, float *sink) {
  float a = seed + threadIdx.x, b = a + 0.3f, c = a + 0.7f, d = a + 1.1f;
  for (int i = 0; i < iters; i += 4) {
    a = __fmul_rn(a, 1.0001f); b = __fmul_rn(b, 1.0001f); c = __fmul_rn(c, 1.0001f); d = __fmul_rn(d, 1.0001f);
  }
  if (a + b + c + d == 123.456f) *sink = a;
}

template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  const int M = 2048;
  std::vector<__half> h((size_t)M * M);
  for (size_t i = 0; i < h.size(); ++i) h[i] = __float2half((float)((i * 2654435761u >> 20) & 255));
  cudaChannelFormatDesc cd = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindFloat);
  cudaArray_t arr; CK(cudaMallocArray(&arr, &cd, M, M, cudaArrayTextureGather));
  CK(cudaMemcpy2DToArray(arr, 0, 0, h.data(), M * 2, M * 2, M, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
  cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType; td.normalizedCoords = 1;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  float *sink; CK(cudaMalloc(&sink, 4));
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int sms = p.multiProcessorCount, iters = 4096;
  printf("%s, %d SMs, clock attr %d MHz\n", p.name, sms, clk_khz / 1000);
  for (int wps : {8, 16, 32, 48}) {     // warps per SM
    const int blocks = sms * wps / 8, threads = 256;
    const double winst = (double)blocks * (threads / 32) * iters;
    for (int mode = 0; mode < 3; ++mode) {
      float ms = timeit([&] { k_tld4<<<blocks, threads>>>(tex, iters, mode == 1 ? 0.9f : 0.05f, 1.0f / M, sink, mode); });
      printf("tld4 mode %d  %2d warps/SM: %.3f ms  %.3f warp-inst/clk/SM (at 1.965 GHz)  %.1f G tld4 lanes/s\n", mode, wps, ms,
             winst / (ms * 1e-3) / 1.965e9 / sms, winst * 32 / (ms * 1e-3) / 1e9);
    }
    float ms = timeit([&] { k_texpoint<<<blocks, threads>>>(tex, iters, 1.0f / M, sink); });
    printf("tex point    %2d warps/SM: %.3f ms  %.3f warp-inst/clk/SM\n", wps, ms, winst / (ms * 1e-3) / 1.965e9 / sms);
    ms = timeit([&] { k_frnd<<<blocks, threads>>>(iters * 4, 1.5f, sink); });
    printf("FRND+FMUL    %2d warps/SM: %.3f ms  %.3f FRND warp-inst/clk/SM\n", wps, ms, winst * 4 / (ms * 1e-3) / 1.965e9 / sms);
    ms = timeit([&] { k_fmul<<<blocks, threads>>>(iters * 4, 1.5f, sink); });
    printf("FMUL         %2d warps/SM: %.3f ms  %.3f FMUL warp-inst/clk/SM\n", wps, ms, winst * 4 / (ms * 1e-3) / 1.965e9 / sms);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
