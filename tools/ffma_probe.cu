// ffma_probe.cu -- developer microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue rate per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma_probe tools/ffma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_scalar(float* out, float a, float b, int iters, long long* clk) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x, y = b;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int NACC>
__global__ void k_packed(float* out, float a, float b, int iters, long long* clk) {
  unsigned long long acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    acc[i] = *reinterpret_cast<unsigned long long*>(&v);
  }
  float2 xv = make_float2(a + threadIdx.x, a - threadIdx.x), yv = make_float2(b, -b);
  unsigned long long x = *reinterpret_cast<unsigned long long*>(&xv), y = *reinterpret_cast<unsigned long long*>(&yv);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(x), "l"(y));
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int iters = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    long long c;
    k_scalar<16><<<148, threads>>>(out, 1.0001f, 0.5f, iters, clk); cudaDeviceSynchronize();
    k_scalar<16><<<148, threads>>>(out, 1.0001f, 0.5f, iters, clk); cudaDeviceSynchronize();
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    double fma_per_clk = (double)threads * 16 * iters / c;
    printf("scalar FFMA  threads %4d: %lld clk, %.1f FMA/clk/SM\n", threads, c, fma_per_clk);
    k_packed<16><<<148, threads>>>(out, 1.0001f, 0.5f, iters, clk); cudaDeviceSynchronize();
    k_packed<16><<<148, threads>>>(out, 1.0001f, 0.5f, iters, clk); cudaDeviceSynchronize();
    cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    fma_per_clk = (double)threads * 16 * 2 * iters / c;
    printf("packed FFMA2 threads %4d: %lld clk, %.1f FMA/clk/SM\n", threads, c, fma_per_clk);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
