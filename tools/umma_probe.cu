// umma_probe.cu -- decodes how tcgen05.mma reads an MN-major SWIZZLE_NONE operand: the probed operand's
// shared memory holds float(word index); the other operand is an identity, so D shows which word the
// hardware fetched for every (mn, k).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tensorbnn_b200/csrc/umma.cuh"
using namespace tbnn;

// which = 0: probe B (MN-major), A = identity K-major [128][8];  D[k][n] = word fetched for B(n,k)
// which = 1: probe A (MN-major), B = identity K-major [N][8];    D[m][k] = word fetched for A(m,k)
constexpr int PW = 40960;   // probed words (160 KB)
__global__ void __launch_bounds__(128, 1) k_probe(float* D, int N, int which, uint32_t lbo, uint32_t sbo, int hi) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* P = reinterpret_cast<float*>(smraw);            // probed operand: 2048 words
  unsigned char* I = smraw + PW * 4;                     // identity operand, K-major core layout
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = tid; i < PW; i += 128) P[i] = hi ? (float)(i / 2048) : (float)(i % 2048);
  const int RI = which != 1 ? 128 : N;                    // rows of the identity operand
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
    const uint64_t dP = umma::smem_desc(smem_u32(P), lbo, sbo);
    if (which == 0) umma::mma_tf32_ss(tbase, dI, dP, umma::idesc_tf32(128, N, false, true), false);
    else if (which == 2) umma::mma_tf32_ss(tbase, dI, dP, umma::idesc_tf32(128, N, false, false), false);
    else umma::mma_tf32_ss(tbase, dP, dI, umma::idesc_tf32(128, N, true, false), false);
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

int main() {
  const int N = 32;
  float* dD; cudaMalloc(&dD, 128 * N * 4);
  std::vector<float> D(128 * N);
  const uint32_t cfgs[][2] = {{128, 1024}, {1024, 128}, {256, 2048}, {2048, 256}, {512, 128}};
  for (int which = 2; which >= 0; --which)
    for (auto& c : cfgs) {
      cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      std::vector<float> Dlo(128 * N);
      cudaMemset(dD, 0xFF, 128 * N * 4);
      k_probe<<<1, 128, 200 * 1024>>>(dD, N, which, c[0], c[1], 0);
      cudaDeviceSynchronize();
      cudaMemcpy(Dlo.data(), dD, Dlo.size() * 4, cudaMemcpyDeviceToHost);
      k_probe<<<1, 128, 200 * 1024>>>(dD, N, which, c[0], c[1], 1);
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      for (size_t i = 0; i < D.size(); ++i) D[i] = D[i] * 2048.f + Dlo[i];
      printf("== which=%d probe %s MN-major  LBO=%u SBO=%u bytes: word index fetched for (mn, k)\n", which, which == 1 ? "A" : "B", c[0], c[1]);
      if (which != 1) {
        for (int k = 0; k < 10; ++k) { printf("k=%d:", k); for (int n = 0; n < N; ++n) printf(" %g", D[k * N + n]); printf("\n"); }
      } else {
        for (int k = 0; k < 8; ++k) { printf("k=%d:", k); for (int m = 0; m < 40; ++m) printf(" %g", D[m * N + k]); printf("\n"); }
      }
    }
  return 0;
}
