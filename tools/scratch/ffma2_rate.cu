// Issue-rate microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 4096
__global__ void k_scalar(float *out, float a, float b) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __fmaf_rn(v[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, float a, float b) {
  unsigned long long v[8], aa, bb;
  float2 t = make_float2(a, a), u = make_float2(b, b);
  aa = *reinterpret_cast<unsigned long long *>(&t);
  bb = *reinterpret_cast<unsigned long long *>(&u);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 x = make_float2(threadIdx.x + 2 * i, threadIdx.x + 2 * i + 1);
    v[i] = *reinterpret_cast<unsigned long long *>(&x);
  }
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(aa), "l"(bb));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 x = *reinterpret_cast<float2 *>(&v[i]);
    s += x.x + x.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float *out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    float ms;
    cudaEventRecord(e0);
    k_scalar<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * 8 * 256 * 16 * ITER;
    printf("scalar FFMA : %.3f ms  %.1f TFLOP/s  %.1f G warp-instr/s\n", ms, 2 * fma / ms / 1e9, fma / 32 / ms / 1e6);
    cudaEventRecord(e0);
    k_packed<<<148 * 8, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("packed FFMA2: %.3f ms  %.1f TFLOP/s  %.1f G warp-instr/s\n", ms, 2 * fma / ms / 1e9, fma / 64 / ms / 1e6);
  }
  return 0;
}
