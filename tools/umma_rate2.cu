// umma_rate2.cu -- what one tcgen05.mma kind::tf32 (M = 128, K = 8) costs when the issuing thread does nothing else:
// the SAME operands issued back to back (no descriptor arithmetic in the loop), descriptors as 32-bit words so that
// they sit in uniform registers.  Separates the hardware rate from the issue overhead tools/umma_rate.cu measures.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate2 tools/umma_rate2.cu
#include <cstdio>
#include <vector>
#include "../tensorbnn_b200/csrc/umma.cuh"
using namespace tbnn;
template <int NMMA>
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int N, int reps, int ndist) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < 48 * 1024; i += 128) reinterpret_cast<float*>(smraw)[i] = 0.f;
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(smraw), b0 = a0 + 64 * 1024;
    const uint32_t id = umma::idesc_tf32(128, N, false, false);
    const uint32_t alo = umma::desc_lo(a0, 2064), blo = umma::desc_lo(b0, 128u * (N / 8) + 16), hi = umma::desc_hi(128u);
    const uint32_t step = ndist ? (2 * 2064) >> 4 : 0;      // ndist: walk through 4 k steps (distinct smem lines) or reuse one
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < NMMA; ++i) umma::mma_tf32_ss32(tbase, alo + (i & 3) * step, hi, blo + (i & 3) * step, hi, id, true);
    }
    umma::commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(k_rate<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int ndist : {0, 1})
    for (int N : {16, 32, 64, 80, 128, 144, 192, 256}) {
      const int reps = 64;
      k_rate<16><<<148, 128, smem>>>(d, N, reps, ndist);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
      printf("%s operands, N=%3d : %.1f cycles / MMA (128 x %d x 8)\n", ndist ? "4 distinct" : "same", N, (double)mx / (reps * 16), N);
    }
  return 0;
}
