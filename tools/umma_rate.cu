// umma_rate.cu -- issue-rate probe for tcgen05.mma kind::tf32 (M = 128, K = 8 per instruction) on a B200:
// cycles per MMA for shared-memory (SS) and tensor-memory (TS) A operands, SWIZZLE_NONE core-matrix layouts with
// either the row groups or the column groups contiguous, several N.  One CTA per SM on every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tensorbnn_b200/csrc/umma.cuh"
using namespace tbnn;

// mode 0: SS, A row groups contiguous (sbo 128, lbo 128*16)     mode 1: SS, A column groups contiguous (lbo 128, sbo = 128*KC/4)
// mode 2: TS (A from TMEM)                                       B always K-major with row groups contiguous
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int N, int mode, int reps, int ksteps, int nacc) {
  extern __shared__ __align__(128) unsigned char smraw[];
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
    const uint32_t a0 = smem_u32(smraw), b0 = a0 + 128 * 1024;
    const uint32_t id = umma::idesc_tf32(128, N, false, false);
    const uint32_t KC = 8 * ksteps;                       // K extent of the staged A tile
    const uint32_t a_lbo = mode != 0 ? 128u : 128u * 16u, a_sbo = mode != 0 ? 32u * KC : 128u;
    const uint32_t a_step = mode != 0 ? 256u : 2u * 128u * 16u;
    const uint32_t bcg = 128u * (N / 8);
    // descriptors are built once; the loop only bumps the 14-bit start-address field (bytes >> 4)
    const uint64_t dA0 = umma::smem_desc(a0, a_lbo, a_sbo);
    const uint64_t dB0 = umma::smem_desc(b0, bcg, 128u);
    const uint32_t am = (uint32_t)nacc - 1u;              // nacc is a power of two
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll 4
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t dA = dA0 + (uint64_t)((ks * a_step) >> 4);
        const uint64_t dB = dB0 + (uint64_t)(((ks & 3) * 2 * bcg) >> 4);
        const uint32_t d = tbase + (uint32_t)((ks & am) * N);     // rotate over nacc independent accumulators
        if (mode == 2) umma::mma_tf32_ts(d, tbase + 384 + 8 * (ks & 15), dB, id, true);
        else if (mode == 3)                                 // descriptors as 32-bit words: they stay in uniform registers
          umma::mma_tf32_ss32(d, umma::desc_lo(a0, a_lbo) + ((ks * a_step) >> 4), umma::desc_hi(a_sbo),
                              umma::desc_lo(b0, bcg) + (((ks & 3) * 2 * bcg) >> 4), umma::desc_hi(128u), id, true);
        else umma::mma_tf32_ss(d, dA, dB, id, true);
      }
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
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int nacc : {1, 2, 4, 8})
    for (int mode = 1; mode < 4; ++mode)
      for (int N : Ns) {
        const int grid = 148;
        if (nacc * N > 384) continue;
        const int reps = 64, ksteps = 16;
        k_rate<<<grid, 128, smem>>>(d, N, mode, reps, ksteps, nacc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("nacc %d mode %d (%s) N=%3d : %.1f cycles / MMA (128 x %d x 8)\n", nacc, mode,
               mode == 0 ? "SS A row-groups contiguous" : mode == 1 ? "SS A col-groups contiguous" : mode == 2 ? "TS A in TMEM" : "SS, 32-bit descriptor words", N,
               (double)mx / (reps * ksteps), N);
      }
  return 0;
}
