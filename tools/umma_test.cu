// umma_test.cu -- standalone validation of the tcgen05 primitives in csrc/umma.cuh on a B200:
// D[128 x N] = A * B for every operand-major combination the BNN tile engine needs, single-pass
// TF32 and error-compensated 3xTF32, operands from shared memory (and A from TMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_test tools/umma_test.cu && ./umma_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tensorbnn_b200/csrc/umma.cuh"

using namespace tbnn;

// mode bits: 1 = A MN-major (A given as [K][M]), 2 = B MN-major (B given as [K][N]), 4 = 3xTF32, 8 = A from TMEM
// A given row-major as Amat[RA][CA]: K-major: [M=128][K]; MN-major: [K][M=128].  Same for B with N.
__global__ void __launch_bounds__(128, 1)
k_umma_test(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K, int mode, int lay) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int M = 128;
  const bool a_mn = mode & 1, b_mn = mode & 2, x3 = mode & 4, a_tm = mode & 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RA = a_mn ? K : M, CA = a_mn ? M : K;
  const int RB = b_mn ? K : N, CB = b_mn ? N : K;
  // core-matrix strides: row groups contiguous (128 B), column groups at 128*(R/8)
  // lay 0: row groups contiguous (128 B), column groups at 128*(R/8); lay 1: column groups contiguous (128 B),
  // row groups at 128*(C/4) (the layout of k_sweep_umma's X chunks)
  const uint32_t rgA = lay ? 128u * (CA / 4) : 128u, cgA = lay ? 128u : 128u * (RA / 8);
  const uint32_t rgB = lay ? 128u * (CB / 4) : 128u, cgB = lay ? 128u : 128u * (RB / 8);
  const uint32_t szA = (uint32_t)RA * CA * 4, szB = (uint32_t)RB * CB * 4;
  unsigned char* sAh = smraw;
  unsigned char* sAl = sAh + szA;
  unsigned char* sBh = sAl + szA;
  unsigned char* sBl = sBh + szB;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int e = tid; e < RA * CA; e += blockDim.x) {
    const int r = e / CA, c = e % CA;
    float hi, lo;
    if (x3) umma::split_tf32(A[e], hi, lo); else { hi = A[e]; lo = 0.f; }
    const uint32_t o = umma::core_off(r, c, rgA, cgA);
    *reinterpret_cast<float*>(sAh + o) = hi;
    *reinterpret_cast<float*>(sAl + o) = lo;
  }
  for (int e = tid; e < RB * CB; e += blockDim.x) {
    const int r = e / CB, c = e % CB;
    float hi, lo;
    if (x3) umma::split_tf32(B[e], hi, lo); else { hi = B[e]; lo = 0.f; }
    const uint32_t o = umma::core_off(r, c, rgB, cgB);
    *reinterpret_cast<float*>(sBh + o) = hi;
    *reinterpret_cast<float*>(sBl + o) = lo;
  }
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tbase = tmem_slot;
  const uint32_t dcol = 0, acol_h = 256, acol_l = 384;   // A (hi / lo) staged in TMEM columns when a_tm (K <= 128)
  if (a_tm) {
    // thread = row of A (K-major only): write this row's K values into TMEM columns
    const int row = tid;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float h[8], l[8];
      for (int i = 0; i < 8; ++i) {
        const float x = A[row * K + k0 + i];
        if (x3) umma::split_tf32(x, h[i], l[i]); else { h[i] = x; l[i] = 0.f; }
      }
      umma::tmem_st8(umma::tmem_addr(tbase, 32 * warp, acol_h + k0), h);
      umma::tmem_st8(umma::tmem_addr(tbase, 32 * warp, acol_l + k0), l);
    }
    umma::tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t id = umma::idesc_tf32(M, N, a_mn, b_mn);
    const uint32_t aH = smem_u32(sAh), aL = smem_u32(sAl), bH = smem_u32(sBh), bL = smem_u32(sBl);
    const uint32_t a_lbo = a_mn ? rgA : cgA, a_sbo = a_mn ? cgA : rgA, a_step = a_mn ? rgA : 2 * cgA;
    const uint32_t b_lbo = b_mn ? rgB : cgB, b_sbo = b_mn ? cgB : rgB, b_step = b_mn ? rgB : 2 * cgB;
    bool acc = false;
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t dAh = umma::smem_desc(aH + ks * a_step, a_lbo, a_sbo);
      const uint64_t dAl = umma::smem_desc(aL + ks * a_step, a_lbo, a_sbo);
      const uint64_t dBh = umma::smem_desc(bH + ks * b_step, b_lbo, b_sbo);
      const uint64_t dBl = umma::smem_desc(bL + ks * b_step, b_lbo, b_sbo);
      const uint32_t d = umma::tmem_addr(tbase, 0, dcol);
      if (a_tm) {
        const uint32_t tAh = umma::tmem_addr(tbase, 0, acol_h + 8 * ks), tAl = umma::tmem_addr(tbase, 0, acol_l + 8 * ks);
        if (x3) {
          umma::mma_tf32_ts(d, tAl, dBh, id, acc); acc = true;
          umma::mma_tf32_ts(d, tAh, dBl, id, acc);
        }
        umma::mma_tf32_ts(d, tAh, dBh, id, acc); acc = true;
      } else {
        if (x3) {
          umma::mma_tf32_ss(d, dAl, dBh, id, acc); acc = true;
          umma::mma_tf32_ss(d, dAh, dBl, id, acc);
        }
        umma::mma_tf32_ss(d, dAh, dBh, id, acc); acc = true;
      }
    }
    umma::commit(&bar);
  }
  mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int n0 = 0; n0 < N; n0 += 8) {
    float v[8];
    umma::tmem_ld8(umma::tmem_addr(tbase, 32 * warp, dcol + n0), v);
    umma::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) D[(32 * warp + lane) * N + n0 + i] = v[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

static float trunc_tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y;
}
static float round_tf32(float x) {
  uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y;
}

int main() {
  const int M = 128;
  int fails = 0;
  const int shapes[][2] = {{64, 64}, {32, 128}, {64, 128}, {64, 8}, {16, 32}, {256, 64}};   // {N, K}
  for (auto& sh : shapes) {
    const int N = sh[0], K = sh[1];
    for (int ml = 0; ml < 32; ++ml) {
      const int mode = ml & 15, lay = ml >> 4;
      if ((mode & 8) && ((mode & 1) || K > 128)) continue;
      std::vector<float> A(M * K), B(N * K), D(M * N, -1.f);
      srand(7 + mode + N);
      for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
      for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
      float *dA, *dB, *dD;
      cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
      cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
      cudaMemset(dD, 0xFF, D.size() * 4);
      const size_t smem = 2 * (size_t)(M * K + N * K) * 4 + 256;
      cudaFuncSetAttribute(k_umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_umma_test<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, lay);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d K=%d mode=%d CUDA error %s\n", N, K, mode, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
      // references: exact (fp64), tf32-truncated inputs, tf32-rounded inputs
      double e_exact = 0, e_trunc = 0, e_round = 0, scale = 0;
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0, st = 0, sr = 0;
          for (int k = 0; k < K; ++k) {
            const float a = (mode & 1) ? A[k * M + m] : A[m * K + k];
            const float b = (mode & 2) ? B[k * N + n] : B[n * K + k];
            s += (double)a * b;
            st += (double)trunc_tf32(a) * trunc_tf32(b);
            sr += (double)round_tf32(a) * round_tf32(b);
          }
          const double d = D[m * N + n];
          e_exact = fmax(e_exact, fabs(d - s)); e_trunc = fmax(e_trunc, fabs(d - st)); e_round = fmax(e_round, fabs(d - sr));
          scale = fmax(scale, fabs(s));
        }
      const bool x3 = mode & 4;
      const double err = x3 ? e_exact : fmin(e_trunc, e_round);
      const bool ok = err <= (x3 ? 2e-6 : 2e-5) * fmax(scale, 1.0);
      printf("lay%d N=%3d K=%3d A:%s%s B:%s %s  err_exact %.3e  err_vs_trunc %.3e  err_vs_round %.3e  %s\n", lay, N, K,
             (mode & 1) ? "MN" : "K ", (mode & 8) ? "(tmem)" : "      ", (mode & 2) ? "MN" : "K ", x3 ? "3xTF32" : "1xTF32",
             e_exact, e_trunc, e_round, ok ? "ok" : "FAIL");
      fails += !ok;
      cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
  }
  printf("%s (%d failures)\n", fails ? "FAILED" : "ALL OK", fails);
  return fails ? 1 : 0;
}
