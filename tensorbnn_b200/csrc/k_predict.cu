// k_predict.cu -- posterior-predictive sweep: stored samples x test rows
// (predictor.predict, predictor.py:132-155).  Each CTA owns a contiguous block of test
// rows and marches through the samples: the sample's padded weights are staged in shared
// memory once per CTA and reused for all of the CTA's rows; the per-row running mean / M2
// (Welford) of the network output live in shared memory across samples, so the fused
// mean/sd mode writes nothing per (sample,row).  The materialising mode writes
// out[S][out][M] coalesced along rows.
#include "engine.cuh"
#include "kernels.h"

namespace tbnn {

template <typename T>
__global__ void __launch_bounds__(NT, 1)
k_predict(const __grid_constant__ ModelPlan mp, const T* __restrict__ samples_pad, long long s0,
          long long S_chunk, const T* __restrict__ X, long long M, int rows_per_cta,
          T* __restrict__ out, T* __restrict__ moments) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* sm = reinterpret_cast<T*>(smraw);
  TileCtx<T> cx;
  cx.sm = sm;
  T* Ws = sm + mp.offW;
  cx.Wp = Ws;
  cx.G = nullptr;
  T* acc = sm + mp.offG;                     // [rows_per_cta][OUT][2] = mean, M2
  const long long r_begin = (long long)blockIdx.x * rows_per_cta;
  if (r_begin >= M) return;
  const int nrows = (int)((M - r_begin) < rows_per_cta ? (M - r_begin) : rows_per_cta);
  const int OUT = mp.OUT, TR = mp.TR;
  const BlockPlan& lb = mp.b[mp.nb - 1];
  if (moments) {
    for (int e = threadIdx.x; e < nrows * OUT; e += blockDim.x) {
      const int lr = e / OUT, o = e - lr * OUT;
      const long long gi = (long long)o * M + r_begin + lr;
      acc[2 * e] = s0 > 0 ? moments[(long long)OUT * M + gi] : T(0);
      acc[2 * e + 1] = s0 > 0 ? moments[2 * (long long)OUT * M + gi] : T(0);
    }
  }
  for (long long s = 0; s < S_chunk; ++s) {
    __syncthreads();
    const T* src = samples_pad + s * mp.Ppad;
    for (int i = 4 * threadIdx.x; i < mp.Ppad; i += 4 * blockDim.x) {
      T v[4];
      ld4(src + i, v);
      st4(Ws + i, v);
    }
    const T cnt = (T)(s0 + s + 1);
    for (int t0 = 0; t0 < nrows; t0 += TR) {
      const int nr = (nrows - t0) < TR ? (nrows - t0) : TR;
      __syncthreads();
      load_x_tile<T>(mp, sm + mp.offX, X, r_begin + t0, nr);
      wait_x_tile();
      __syncthreads();
      tile_forward<T>(mp, cx);
      const T* F = sm + lb.offS;
      for (int e = threadIdx.x; e < nr * OUT; e += blockDim.x) {
        const int o = e / nr, r = e - o * nr;              // rows fastest: coalesced stores
        const T f = F[r * lb.ld_out + o];
        if (out) out[((s0 + s) * OUT + o) * M + r_begin + t0 + r] = f;
        if (moments) {
          const int a = 2 * ((t0 + r) * OUT + o);
          const T mean = acc[a], d = f - mean;
          const T mnew = mean + d / cnt;
          acc[a] = mnew;
          acc[a + 1] += d * (f - mnew);
        }
      }
    }
  }
  __syncthreads();
  if (moments) {
    for (int e = threadIdx.x; e < nrows * OUT; e += blockDim.x) {
      const int lr = e / OUT, o = e - lr * OUT;
      const long long gi = (long long)o * M + r_begin + lr;
      moments[gi] = (T)(s0 + S_chunk);
      moments[(long long)OUT * M + gi] = acc[2 * e];
      moments[2 * (long long)OUT * M + gi] = acc[2 * e + 1];
    }
  }
}

template <typename T>
void Launch<T>::predict(const ModelPlan& mp, const T* samples_pad, long long s0, long long S_chunk,
                        long long S_total, const T* X, long long M, int rows_per_cta, T* out,
                        T* moments, cudaStream_t st) {
  (void)S_total;
  const int grid = (int)((M + rows_per_cta - 1) / rows_per_cta);
  const size_t smem = (size_t)mp.smem_elems * sizeof(T);
  cudaFuncSetAttribute(k_predict<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_predict<T><<<grid, NT, smem, st>>>(mp, samples_pad, s0, S_chunk, X, M, rows_per_cta, out, moments);
}

template struct Launch<float>;
template struct Launch<double>;

}  // namespace tbnn
