// Microbenchmark: how fast can tiles of TC columns x TR rows be written into row-major frames?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int TC, int TR>
__global__ void __launch_bounds__(256) k(uint32_t* out, int w, int h, size_t pose_stride) {
  const int c0 = blockIdx.x * TC, r0 = blockIdx.y * TR;
  uint32_t* o = out + (size_t)blockIdx.z * pose_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SEGS = TC / 32;                 // 128-byte segments per tile row
  for (int i = warp; i < TR * SEGS; i += 8) {
    const int r = r0 + i / SEGS, c = c0 + (i % SEGS) * 32 + lane;
    if (r < h && c < w) o[(size_t)r * w + c] = 0xFF000000u | (r ^ c);
  }
}
template <int TC, int TR>
void run(uint32_t* d, int w, int h, int poses) {
  dim3 grid((w + TC - 1) / TC, (h + TR - 1) / TR, poses);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) k<TC, TR><<<grid, 256>>>(d, w, h, (size_t)w * h);
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) k<TC, TR><<<grid, 256>>>(d, w, h, (size_t)w * h);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("tile %4d x %4d: %.1f us/frame  %.0f GB/s\n", TC, TR, ms / 10 / poses * 1e3, (double)w * h * 4 * poses * 10 / ms / 1e6);
}
int main() {
  const int w = 1920, h = 1080, poses = 48;
  uint32_t* d; cudaMalloc(&d, (size_t)w * h * 4 * poses);
  run<32, 256>(d, w, h, poses);
  run<64, 128>(d, w, h, poses);
  run<128, 64>(d, w, h, poses);
  run<256, 32>(d, w, h, poses);
  run<1920, 8>(d, w, h, poses);
  cudaMemset(d, 0, (size_t)w * h * 4 * poses);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); for (int i = 0; i < 10; ++i) cudaMemsetAsync(d, i, (size_t)w * h * 4 * poses); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); printf("memset: %.0f GB/s\n", (double)w * h * 4 * poses * 10 / ms / 1e6);
  return 0;
}
