// k_adapter.cu -- paramAdapter.gridSearch on device (paramAdapter.py:158-196): one thread
// per (step size, leapfrog count) grid point evaluates the UCB of calcUCB (:113-141) with
// the product kernel of calck (:95-111, Q11); a two-stage arg-max keeps the reference's
// "first maximum in scan order" (e fastest, L slowest, strict '>').  float32 throughout,
// like the reference (:60).
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.h"

namespace tbnn {

constexpr int AH = 64;   // max history length (reference caps at 50, paramAdapter.py:285-289)

struct UcbArgs {
  int eN, lN, n;
  float s, p, rootbeta, el, eu, Ll, Lu, s00, s01, s10, s11;
};

__device__ __forceinline__ void better(float& v, int& i, float v2, int i2) {
  if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
}

__global__ void __launch_bounds__(256)
k_ucb(UcbArgs a, const float* __restrict__ eGrid, const float* __restrict__ lGrid,
      const float* __restrict__ prev, const float* __restrict__ Kinv, const float* __restrict__ KinvR,
      float* __restrict__ blk_val, int* __restrict__ blk_idx) {
  __shared__ float pg[AH][2], sKinv[AH * AH], sKR[AH];
  __shared__ float rv[8];
  __shared__ int ri[8];
  for (int j = threadIdx.x; j < a.n; j += blockDim.x) {
    pg[j][0] = -1.0f + 2.0f * (prev[2 * j] - a.el) / (a.eu - a.el);
    pg[j][1] = -1.0f + 2.0f * (prev[2 * j + 1] - a.Ll) / (a.Lu - a.Ll);
    sKR[j] = KinvR[j];
  }
  for (int j = threadIdx.x; j < a.n * a.n; j += blockDim.x) sKinv[j] = Kinv[j];
  __syncthreads();
  const int total = a.eN * a.lN;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float val = -1000000000.0f;
  int best = 0x7fffffff;
  if (idx < total) {
    const int lc = idx / a.eN, ec = idx - lc * a.eN;
    const float ge = -1.0f + 2.0f * (eGrid[ec] - a.el) / (a.eu - a.el);
    const float gl = -1.0f + 2.0f * (lGrid[lc] - a.Ll) / (a.Lu - a.Ll);
    // Sigma * gamma_test
    const float t0 = a.s00 * ge + a.s01 * gl, t1 = a.s10 * ge + a.s11 * gl;
    float kv[AH];
    float mean = 0.0f;
    for (int j = 0; j < a.n; ++j) {
      kv[j] = expf(-0.5f * (pg[j][0] * t0 + pg[j][1] * t1));
      mean += kv[j] * sKR[j];
    }
    mean *= a.s;
    float quad = 0.0f;
    for (int i = 0; i < a.n; ++i) {
      float r = 0.0f;
      for (int j = 0; j < a.n; ++j) r += sKinv[i * a.n + j] * kv[j];
      quad += kv[i] * r;
    }
    const float kself = expf(-0.5f * (ge * t0 + gl * t1));
    const float ucb = mean + (kself - quad) * a.p * a.rootbeta;
    if (ucb > val) { val = ucb; best = idx; }
  }
  // block arg-max (max value, then min index)
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, val, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best, o);
    better(val, best, v2, i2);
  }
  if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = val; ri[threadIdx.x >> 5] = best; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) better(val, best, rv[w], ri[w]);
    blk_val[blockIdx.x] = val;
    blk_idx[blockIdx.x] = best;
  }
}

__global__ void k_ucb_final(UcbArgs a, int nblk, const float* __restrict__ eGrid,
                            const float* __restrict__ lGrid, const float* __restrict__ blk_val,
                            const int* __restrict__ blk_idx, float* __restrict__ out) {
  __shared__ float rv[8];
  __shared__ int ri[8];
  float val = -1000000000.0f;
  int best = 0x7fffffff;
  for (int j = threadIdx.x; j < nblk; j += blockDim.x) better(val, best, blk_val[j], blk_idx[j]);
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, val, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, best, o);
    better(val, best, v2, i2);
  }
  if ((threadIdx.x & 31) == 0) { rv[threadIdx.x >> 5] = val; ri[threadIdx.x >> 5] = best; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) better(val, best, rv[w], ri[w]);
    if (best == 0x7fffffff || !(val > -1000000000.0f)) {   // nothing beat the start value
      out[0] = a.el; out[1] = a.Ll; out[2] = -1000000000.0f;
    } else {
      const int lc = best / a.eN, ec = best - lc * a.eN;
      out[0] = eGrid[ec]; out[1] = lGrid[lc]; out[2] = val;
    }
  }
}

size_t adapter_workspace_bytes(int eNumber, int lNumber) {
  const size_t total = (size_t)eNumber * lNumber, nblk = (total + 255) / 256;
  return nblk * (sizeof(float) + sizeof(int)) + 64;
}

void launch_adapter_ucb(const float* eGrid, int eNumber, const float* lGrid, int lNumber,
                        const float* prev, int n_hist, const float* Kinv, const float* KinvR, float s,
                        float p, float rootbeta, float el, float eu, float Ll, float Lu,
                        const float* sigma, float* out, void* workspace, cudaStream_t st) {
  UcbArgs a;
  a.eN = eNumber; a.lN = lNumber; a.n = n_hist;
  a.s = s; a.p = p; a.rootbeta = rootbeta; a.el = el; a.eu = eu; a.Ll = Ll; a.Lu = Lu;
  a.s00 = sigma[0]; a.s01 = sigma[1]; a.s10 = sigma[2]; a.s11 = sigma[3];
  const int total = eNumber * lNumber, nblk = (total + 255) / 256;
  float* bv = reinterpret_cast<float*>(workspace);
  int* bi = reinterpret_cast<int*>(bv + nblk);
  k_ucb<<<nblk, 256, 0, st>>>(a, eGrid, lGrid, prev, Kinv, KinvR, bv, bi);
  k_ucb_final<<<1, 256, 0, st>>>(a, nblk, eGrid, lGrid, bv, bi, out);
}

}  // namespace tbnn
