// umma_probe_sw.cu -- decodes how tcgen05.mma kind::tf32 reads an MN-major operand in the SWIZZLE_128B_BASE32B
// layout (descriptor layout type 1; MN-major SWIZZLE_NONE returns zeros for 32-bit operands, tools/umma_probe.cu).
// The probed operand's shared memory holds float(word index); the other operand is a K-major identity, so D shows
// which word the hardware fetched for every (mn, k).  The fetched addresses are compared with the hypothesis
//   byte(mn, k) = (mn / 32) * LBO + (k / 4) * SBO + (k % 4) * 128 + (((mn % 32) / 8) ^ (k % 4)) * 32 + (mn % 8) * 4
// and with LBO / SBO exchanged.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe_sw tools/umma_probe_sw.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tensorbnn_b200/csrc/umma.cuh"
using namespace tbnn;

constexpr int PW = 40960;   // probed words (160 KB)
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return umma::smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}
__global__ void __launch_bounds__(128, 1) k_probe(float* D, int N, int which, uint32_t lbo, uint32_t sbo, int hi, uint32_t layout,
                                                  uint32_t start_off) {
  extern __shared__ __align__(1024) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* P = reinterpret_cast<float*>(smraw);            // probed operand
  unsigned char* I = smraw + PW * 4;                     // identity operand, K-major SWIZZLE_NONE core layout
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < PW; i += 128) P[i] = hi ? (float)(i / 2048) : (float)(i % 2048);
  const int RI = which == 0 ? 128 : N;                   // rows of the identity operand
  for (int e = tid; e < RI * 8; e += 128) {
    const int r = e / 8, c = e % 8;
    *reinterpret_cast<float*>(I + umma::core_off(r, c, 128, 128u * (RI / 8))) = (r == c) ? 1.f : 0.f;
  }
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  if (tid == 0) {
    const uint64_t dI = umma::smem_desc(smem_u32(I), 128u * (RI / 8), 128);
    const uint64_t dP = desc_sw(smem_u32(P) + start_off, lbo, sbo, layout);
    if (which == 0) umma::mma_tf32_ss(tbase, dI, dP, umma::idesc_tf32(128, N, false, true), false);   // probe B: D[k][n]
    else umma::mma_tf32_ss(tbase, dP, dI, umma::idesc_tf32(128, N, true, false), false);              // probe A: D[m][k]
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    umma::tmem_ld8(umma::tmem_addr(tbase, 32 * warp, n0), v);
    umma::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) D[(32 * warp + lane) * N + n0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 256);
}

static long hyp(int mn, int k, long lbo, long sbo) {
  return (mn / 32) * lbo + (k / 4) * sbo + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4;
}

int main() {
  float* dD; cudaMalloc(&dD, 128 * 128 * 4);
  const uint32_t cfgs[][3] = {{4096, 512, 0}, {512, 4096, 0}, {8192, 1024, 0}, {1024, 8192, 0}, {4096, 512, 1024}, {4096, 512, 128}};
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int which = 0; which < 2; ++which) {
    const int N = which == 0 ? 64 : 16;                  // probe B: N = 64 (two 32-wide blocks); probe A: identity B [16][8]
    const int MN = which == 0 ? N : 128;
    for (auto& c : cfgs) {
      std::vector<float> Dlo(128 * N), D(128 * N);
      cudaMemset(dD, 0xFF, 128 * N * 4);
      k_probe<<<1, 128, 200 * 1024>>>(dD, N, which, c[0], c[1], 0, 1, c[2]);
      cudaDeviceSynchronize();
      cudaMemcpy(Dlo.data(), dD, Dlo.size() * 4, cudaMemcpyDeviceToHost);
      k_probe<<<1, 128, 200 * 1024>>>(dD, N, which, c[0], c[1], 1, 1, c[2]);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < D.size(); ++i) D[i] = D[i] * 2048.f + Dlo[i];
      int bad1 = 0, bad2 = 0;
      for (int k = 0; k < 8; ++k)
        for (int mn = 0; mn < MN; ++mn) {
          const long got = 4 * (long)(which == 0 ? D[k * N + mn] : D[mn * N + k]) - c[2];
          if (got != hyp(mn, k, c[0], c[1])) ++bad1;
          if (got != hyp(mn, k, c[1], c[0])) ++bad2;
        }
      printf("== probe %s MN-major SW128_32B LBO=%u SBO=%u start+%u: mismatches vs hypothesis %d, with LBO/SBO exchanged %d\n",
             which == 0 ? "B" : "A", c[0], c[1], c[2], bad1, bad2);
      for (int k = 0; k < 8; ++k) {
        printf("k=%d:", k);
        for (int mn = 0; mn < (MN < 72 ? MN : 72); ++mn) printf(" %ld", 4 * (long)(which == 0 ? D[k * N + mn] : D[mn * N + k]));
        printf("\n");
      }
    }
  }
  return 0;
}
