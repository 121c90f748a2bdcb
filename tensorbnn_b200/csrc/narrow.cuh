// narrow.cuh -- the "narrow tail" of a network: dense blocks l >= 1 whose widths are <= 32
// (e.g. 20 -> 20 -> 1 behind a 784-wide first layer, or the whole 1-10-10-10-1 example net).
//
// One training row per HALF-WARP, lane owns two neurons: forward (layer.py:266-279 + activation), likelihood
// residual (likelihood.py:88-94,162-167,225-236) and the data gradient back to dz of block 0 run
// without any CTA-wide barrier (only __syncwarp).  Per-row activations a_l and dz_l are left in
// batch buffers in shared memory ([RB rows][ld]); the weight / bias / slope gradients of blocks
// >= 1 are accumulated from those buffers once per batch (narrow_accum), a small tile GEMM over rows.
#pragma once
#include "engine.cuh"

namespace tbnn {

// One row per HALF-WARP: lane = (half, j), the lane owns neurons j and j + 16 of every block.
// `rb` is this half-warp's row in the batch buffers (active = false: the row does not exist, the
// lanes only take part in the warp barriers).  In: S_0 row (and Z_0 row when block 0 keeps z)
// filled by the caller.  Out: dz of block 0 in dz0_row[0..out_p0) (and the slope contribution c_0
// in the Z_0 row); returns this lane's contribution to the likelihood statistic.
template <typename T>
__device__ __forceinline__ T narrow_row(const ModelPlan& mp, const T* Wp, T* sm, int rb, int j, bool active,
                                        const T* __restrict__ Y, long long row, T* dz0_row) {
  const int OUT = mp.OUT, nb = mp.nb;
  T yv[2] = {T(0), T(0)};
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (active && j + 16 * i < OUT) yv[i] = Y[row * (long long)OUT + j + 16 * i];
  T a[2] = {T(0), T(0)}, z[2] = {T(0), T(0)};
  if (nb == 1 && active) {
    const BlockPlan& b = mp.b[0];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int o = j + 16 * i;
      if (o < b.out_p) {
        a[i] = sm[b.offS + rb * b.ld_out + o];
        if (b.offZ >= 0) z[i] = sm[b.offZ + rb * b.ld_out + o];
      }
    }
  }
  // ---- forward through blocks 1..nb-1
  for (int l = 1; l < nb; ++l) {
    const BlockPlan& b = mp.b[l];
    if (active) {
      const T* ap = sm + mp.b[l - 1].offS + rb * b.ld_in;
      const int kch = b.in_p >> 2;
      const bool two = j + 16 < b.out_p;
      const T* w0 = Wp + b.pw + j * b.ld_in;
      const T* w1 = w0 + 16 * b.ld_in;
      T s[2][2] = {{T(0), T(0)}, {T(0), T(0)}};
      if (j < b.out_p) {
        for (int kc = 0; kc < kch; ++kc) {
          T av[4], wv[4];
          ld4(ap + 4 * kc, av);
          ld4(w0 + 4 * kc, wv);
          s[0][0] = fma(wv[0], av[0], s[0][0]); s[0][1] = fma(wv[1], av[1], s[0][1]);
          s[0][0] = fma(wv[2], av[2], s[0][0]); s[0][1] = fma(wv[3], av[3], s[0][1]);
          if (two) {
            ld4(w1 + 4 * kc, wv);
            s[1][0] = fma(wv[0], av[0], s[1][0]); s[1][1] = fma(wv[1], av[1], s[1][1]);
            s[1][0] = fma(wv[2], av[2], s[1][0]); s[1][1] = fma(wv[3], av[3], s[1][1]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int o = j + 16 * i;
        a[i] = T(0); z[i] = T(0);
        if (o < b.out_p) {
          if (o < b.out) {
            z[i] = (s[i][0] + s[i][1]) + Wp[b.pb + o];
            T slope = T(0);
            if (act_keeps_z(b.act)) slope = eff_slope<T>(b.act, Wp + (b.ps >= 0 ? b.ps : 0), o, T(b.alpha));
            a[i] = act_fwd<T>(b.act, z[i], slope);
          }
          sm[b.offS + rb * b.ld_out + o] = a[i];
          if (b.offZ >= 0) sm[b.offZ + rb * b.ld_out + o] = z[i];
        }
      }
    }
    __syncwarp();
  }
  // ---- likelihood residual -> dz of the last block (same arithmetic as lik_phase)
  T stat = T(0);
  {
    const BlockPlan& b = mp.b[nb - 1];
    if (active) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int o = j + 16 * i;
        if (o >= b.out_p) continue;
        T dz = T(0), c = T(0);
        if (o < OUT) {
          const T f = a[i], y = yv[i];
          T df;
          if (mp.lik == LIK_BERN) {
            const T lo = T(1e-8), hi = T(1 - 1e-7);
            const T p = f < lo ? lo : (f > hi ? hi : f);
            stat += (T(1) - y) * t_log1p(-p) + y * t_log(p);
            df = (f < lo || f > hi) ? T(0) : (y / p - (T(1) - y) / (T(1) - p));
          } else {
            const T res = y - f;
            stat = fma(res, res, stat);
            df = res;
          }
          if (act_keeps_z(b.act)) {
            const bool neg = z[i] < T(0);
            const T sl = eff_slope<T>(b.act, Wp + (b.ps >= 0 ? b.ps : 0), o, T(b.alpha));
            dz = neg ? df * sl : df;
            c = neg ? z[i] * df : T(0);
          } else {
            dz = df * act_deriv_from_out<T>(b.act, f);
          }
        }
        T* dst = nb == 1 ? dz0_row : sm + b.offD + rb * b.ld_out;
        dst[o] = dz;
        if (act_has_slopes(b.act)) sm[b.offZ + rb * b.ld_out + o] = c;
      }
    }
    __syncwarp();
  }
  // ---- data gradient: dz_{l-1}[k] = (sum_o W_l[o][k] dz_l[o]) * act'_{l-1}
  for (int l = nb - 1; l >= 1; --l) {
    const BlockPlan& b = mp.b[l];
    const BlockPlan& pb = mp.b[l - 1];
    if (active && j < b.in_p) {
      const bool two = j + 16 < b.in_p;
      const T* wc0 = Wp + b.pw + j;
      const T* dzr = sm + b.offD + rb * b.ld_out;
      const int ld = b.ld_in, och = b.out_p >> 2;
      T s[2][2] = {{T(0), T(0)}, {T(0), T(0)}};
      for (int oc = 0; oc < och; ++oc) {
        T dv[4];
        ld4(dzr + 4 * oc, dv);
        const T* wc = wc0 + (4 * oc) * ld;
        s[0][0] = fma(dv[0], wc[0], s[0][0]); s[0][1] = fma(dv[1], wc[ld], s[0][1]);
        s[0][0] = fma(dv[2], wc[2 * ld], s[0][0]); s[0][1] = fma(dv[3], wc[3 * ld], s[0][1]);
        if (two) {
          s[1][0] = fma(dv[0], wc[16], s[1][0]); s[1][1] = fma(dv[1], wc[ld + 16], s[1][1]);
          s[1][0] = fma(dv[2], wc[2 * ld + 16], s[1][0]); s[1][1] = fma(dv[3], wc[3 * ld + 16], s[1][1]);
        }
      }
      T* dst = l == 1 ? dz0_row : sm + pb.offD + rb * pb.ld_out;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = j + 16 * i;
        if (k >= b.in_p) continue;
        const T da = s[i][0] + s[i][1];
        T dzp = T(0), cp = T(0);
        if (k < pb.out) {
          if (act_keeps_z(pb.act)) {
            const T zz = sm[pb.offZ + rb * pb.ld_out + k];
            const bool neg = zz < T(0);
            const T sl = eff_slope<T>(pb.act, Wp + (pb.ps >= 0 ? pb.ps : 0), k, T(pb.alpha));
            dzp = neg ? da * sl : da;
            cp = neg ? zz * da : T(0);
          } else {
            dzp = da * act_deriv_from_out<T>(pb.act, sm[pb.offS + rb * pb.ld_out + k]);
          }
        }
        dst[k] = dzp;
        if (act_has_slopes(pb.act)) sm[pb.offZ + rb * pb.ld_out + k] = cp;
      }
    }
    __syncwarp();
  }
  return stat;
}

// Gradient accumulation of blocks 1..nb-1 over the first `nrows` rows of the batch buffers:
//   G.W_l += dZ_l^T S_{l-1},  G.b_l += colsum dZ_l,  slope gradients from the c values left in Z_l.
// Every (l, o, k) is owned by exactly one of the `nthr` cooperating threads (tid = 0..nthr-1); the caller
// synchronises them before and after.
template <typename T>
__device__ __forceinline__ void narrow_accum(const ModelPlan& mp, const T* Wp, T* G, const T* sm, int nrows,
                                             int tid, int nthr) {
  int total = 0;
  for (int l = 1; l < mp.nb; ++l) total += (mp.b[l].out_p >> 2) * (mp.b[l].in_p >> 2);
  for (int t = tid; t < total; t += nthr) {
    int l = 1, u = t;
    for (; l < mp.nb; ++l) {
      const int n = (mp.b[l].out_p >> 2) * (mp.b[l].in_p >> 2);
      if (u < n) break;
      u -= n;
    }
    const BlockPlan& b = mp.b[l];
    const int tn = b.out_p >> 2, og = u % tn, kg = u / tn;
    const T* dz = sm + b.offD + 4 * og;
    const T* ar = sm + mp.b[l - 1].offS + 4 * kg;
    const int ldz = b.ld_out, lda = b.ld_in;
    T acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = T(0);
#pragma unroll 2
    for (int r = 0; r < nrows; ++r) {
      T dv[4], av[4];
      ld4(dz + r * ldz, dv);
      ld4(ar + r * lda, av);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = fma(dv[j], av[q], acc[j][q]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T g[4];
      T* gp = G + b.pw + (4 * og + j) * lda + 4 * kg;
      ld4(gp, g);
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] += acc[j][q];
      st4(gp, g);
    }
  }
  // column sums, handed out from the top of the CTA so they overlap with the tiles above
  int ncol = 0;
  for (int l = 1; l < mp.nb; ++l) ncol += mp.b[l].out_p * (act_has_slopes(mp.b[l].act) ? 2 : 1);
  for (int t = nthr - 1 - tid; t < ncol; t += nthr) {
    int l = 1, u = t;
    for (; l < mp.nb; ++l) {
      const int n = mp.b[l].out_p * (act_has_slopes(mp.b[l].act) ? 2 : 1);
      if (u < n) break;
      u -= n;
    }
    const BlockPlan& b = mp.b[l];
    const bool slope = u >= b.out_p;
    const int o = slope ? u - b.out_p : u;
    const T* src = sm + (slope ? b.offZ : b.offD) + o;
    T s0 = T(0), s1 = T(0);
    int r = 0;
    for (; r + 1 < nrows; r += 2) { s0 += src[r * b.ld_out]; s1 += src[(r + 1) * b.ld_out]; }
    if (r < nrows) s0 += src[r * b.ld_out];
    const T s = s0 + s1;
    if (slope) {
      const T f = b.act == ACT_SQPRELU ? T(2) * Wp[b.ps + o] : T(1);
      G[b.ps + o] += f * s;
    } else {
      G[b.pb + o] += s;
    }
  }
}

template <typename T>
__device__ __forceinline__ void narrow_accum(const ModelPlan& mp, const T* Wp, T* G, const T* sm, int nrows) {
  narrow_accum<T>(mp, Wp, G, sm, nrows, (int)threadIdx.x, (int)blockDim.x);
}

}  // namespace tbnn
