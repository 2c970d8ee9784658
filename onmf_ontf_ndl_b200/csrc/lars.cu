// K3 -- batched positive LARS-lasso sparse coder for sm_100a.
//
// Replaces SparseCoder(transform_algorithm='lasso_lars', positive_code=True).transform as called at
// reference src/ontf.py:79-86, i.e. the per-sample Python loop of LassoLars._fit
// (sklearn/linear_model/_least_angle.py:1136-1153) around _lars_path_solver (:415-917; Gram mode,
// method='lasso', positive=True, return_path=False).
//
// Mapping: one lane-group (LPC = 8/16/32 lanes) follows the homotopy path of one minibatch column; a
// warp carries 32/LPC columns, a CTA carries NW warps, columns are handed out through a global ticket
// counter.  The Gram matrix G = W^T W is staged once per CTA in shared memory (when it fits); each
// group keeps the inverse of the active Gram block, M = G_AA^-1, in its own shared-memory tile and
// maintains it by bordering (atom joins) / Schur downdate (atom leaves).  That replaces sklearn's
// Cholesky factor + two triangular solves per knot (3 s dependent steps) by two s x s lane-parallel
// passes; in fp32 one step of iterative refinement on the equiangular weights restores the accuracy
// (measured: 2.6e-5 rel. code error on a cond(G)=1.5e5 learned dictionary versus 1.2e-3 without).
// All lane<->lane traffic is warp shuffles; the only barriers are __syncwarp().
//
// Path semantics reproduced from sklearn (so the result matches the reference also where sklearn is
// not at the exact lasso optimum):
//   - join: inactive atom with the largest covariance (ties: lowest index)
//   - recorded alpha of a knot = max INACTIVE covariance / d; stop when alpha <= alpha/d + eps32 and
//     interpolate linearly between the last two coefficient vectors
//   - step gamma = min(min_pos((C-c_i)/(AA-a_i+tiny32)), C/AA); drop when a coefficient would cross 0
//     first (gamma = z_pos), no atom joins on the iteration after a drop, the dropped atom's covariance
//     is recomputed exactly
//   - "alpha increasing" bail-out, degenerate-pivot rejection (cov := 0), max_iter.
#include <math_constants.h>

#include "common.cuh"

namespace onmf {

template <typename T>
struct LarsParams {
  const T* G;        // k x k
  const T* Ct;       // n x k
  T* Ht;             // n x k
  long long n;
  int k, d, max_iter;
  T amin;            // alpha / d
  unsigned long long* ticket;        // work counter (zeroed by the host wrapper)
  const long long* col_list;         // overflow pass: columns to solve (else nullptr)
  const unsigned int* n_list;        // overflow pass: device-side count
  long long* ovf_list;               // main pass: columns whose active set outgrew SMAX
  unsigned int* ovf_count;
  T* Mscratch;                       // overflow pass: per-group M storage in global memory
  onmf_lars_stats* stats;
};

template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
  static __device__ __forceinline__ float big() { return 3.402823466e+38f; }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double inf() { return CUDART_INF; }
  static __device__ __forceinline__ double big() { return 1.7976931348623157e+308; }
};

template <typename T, int LPC>
__device__ __forceinline__ T gsum(T v) {
#pragma unroll
  for (int off = LPC / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
template <typename T, int LPC>
__device__ __forceinline__ T gmin(T v) {
#pragma unroll
  for (int off = LPC / 2; off > 0; off >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o < v ? o : v;
  }
  return v;
}
template <int LPC>
__device__ __forceinline__ int gmini(int v) {
#pragma unroll
  for (int off = LPC / 2; off > 0; off >>= 1) {
    int o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o < v ? o : v;
  }
  return v;
}
template <int LPC>
__device__ __forceinline__ int gmaxi(int v) {
#pragma unroll
  for (int off = LPC / 2; off > 0; off >>= 1) {
    int o = __shfl_xor_sync(0xffffffffu, v, off);
    v = o > v ? o : v;
  }
  return v;
}

// shared-memory words (4 B) one group needs, padded so that consecutive groups of a warp start LPC banks apart
template <typename T, int LPC, int SMAX, bool MGLOB>
__host__ __device__ constexpr int group_words() {
  int tw = sizeof(T) / 4;
  int w = (MGLOB ? 0 : SMAX * SMAX * tw) + 3 * SMAX * tw + SMAX;
  if (LPC < 32) {
    int r = w % 32;
    int want = LPC % 32;
    w += (want - r + 32) % 32;
  }
  return w;
}
template <int LPC, int NA>
__host__ __device__ constexpr int gram_stride() {
  return LPC * NA + (LPC < 32 ? LPC : 0);
}

template <typename T, int LPC, int NA, int SMAX, bool GSM, bool MGLOB>
__global__ void __launch_bounds__(512, 1) lars_kernel(LarsParams<T> P) {
  constexpr int SA = SMAX / LPC;       // active slots per lane
  constexpr int GPW = 32 / LPC;        // columns per warp
  constexpr int GS = gram_stride<LPC, NA>();
  constexpr bool REFINE = (sizeof(T) == 4);
  static_assert(SMAX % LPC == 0, "SMAX must be a multiple of LPC");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k;
  T* Gs = reinterpret_cast<T*>(smem_raw);
  const size_t g_bytes = GSM ? round_up<size_t>((size_t)k * GS * sizeof(T), 128) : 0;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int l = lane % LPC;
  const int gid = warp * GPW + lane / LPC;           // group id inside the CTA
  uint32_t* gbase = reinterpret_cast<uint32_t*>(smem_raw + g_bytes) +
                    (size_t)gid * group_words<T, LPC, SMAX, MGLOB>();
  T* Mg;
  T* vecs;
  if (MGLOB) {
    Mg = P.Mscratch + ((size_t)blockIdx.x * (blockDim.x / LPC) + gid) * (size_t)SMAX * SMAX;
    vecs = reinterpret_cast<T*>(gbase);
  } else {
    Mg = reinterpret_cast<T*>(gbase);
    vecs = Mg + SMAX * SMAX;
  }
  T* gs = vecs;                 // g = G[active, j]   (also the refinement residual)
  T* us = vecs + SMAX;          // u = M g            (also the dropped row of M)
  T* ws = vecs + 2 * SMAX;      // equiangular weights by slot
  int* acts = reinterpret_cast<int*>(vecs + 3 * SMAX);   // slot -> atom (-1 = free)

  if (GSM) {
    for (int idx = threadIdx.x; idx < k * GS; idx += blockDim.x) {
      int a = idx / GS, i = idx - a * GS;
      Gs[idx] = (i < k) ? P.G[(size_t)a * k + i] : T(0);
    }
    __syncthreads();
  }
  auto Gat = [&](int a, int i) -> T {
    if (GSM) return Gs[a * GS + i];
    return (i < k) ? __ldg(P.G + (size_t)a * k + i) : T(0);
  };

  const T tiny = T(1.1754943508222875e-38);      // np.finfo(np.float32).tiny
  const T eps32 = T(1.1920928955078125e-07);     // np.finfo(np.float32).eps  (equality_tolerance)
  const T piv_floor = T(2.220446049250313e-16);  // LassoLars eps default
  const T dT = T(P.d);
  const T amin = P.amin;

  unsigned long long st_knots = 0, st_s = 0, st_s2 = 0, st_drops = 0, st_flag = 0, st_cols = 0, st_ovf = 0;
  int st_maxact = 0;

  while (true) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(P.ticket, (unsigned long long)GPW);
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned long long nwork = P.col_list ? (unsigned long long)(*P.n_list) : (unsigned long long)P.n;
    if (base >= nwork) break;
    const unsigned long long widx = base + lane / LPC;
    const bool valid = widx < nwork;
    const long long col = valid ? (P.col_list ? P.col_list[widx] : (long long)widx) : 0;
    const T* crow = P.Ct + (size_t)col * k;

    // ---- per-column state ----
    T cov[NA];
    unsigned inact = 0;
#pragma unroll
    for (int m = 0; m < NA; ++m) {
      int i = l + LPC * m;
      bool ok = valid && i < k;
      cov[m] = ok ? crow[i] : T(0);
      if (ok) inact |= 1u << m;
    }
    T coef[SA], prev[SA];
#pragma unroll
    for (int m = 0; m < SA; ++m) { coef[m] = T(0); prev[m] = T(0); }
    int n_iter = 0, n_act = 0, hw = 0, status = 0, max_act = 0;
    bool drop = false, done = !valid;
    int dslot = 0;
    T a_prev = T(0);
    // the atom dropped by the last step: it is inactive from now on, but sklearn's prev_coef still holds
    // its value at the previous knot, which matters when the path stops inside the segment that ended with
    // the drop (the interpolation then lands on a point where the atom is still positive).
    int ghost_atom = -1;
    T ghost_prev = T(0), ghost_val = T(0);

    while (!__all_sync(0xffffffffu, done)) {
      __syncwarp();
      // ---- 1. largest inactive covariance ----
      T best = -Num<T>::inf();
      int bi = 0x7fffffff;
#pragma unroll
      for (int m = 0; m < NA; ++m)
        if ((inact >> m) & 1u) {
          if (cov[m] > best) { best = cov[m]; bi = l + LPC * m; }
        }
#pragma unroll
      for (int off = LPC / 2; off > 0; off >>= 1) {
        T ov = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      const bool any_inact = (bi != 0x7fffffff);
      const T C = any_inact ? best : T(0);
      const T a_cur = C / dT;
      bool do_add = false, skip = false;
      if (!done) {
        if (a_cur <= amin + eps32) {
          T diff = a_cur - amin;
          if ((diff > eps32 || diff < -eps32) && n_iter > 0) {
            T ss = (a_prev - amin) / (a_prev - a_cur);
#pragma unroll
            for (int m = 0; m < SA; ++m) coef[m] = prev[m] + ss * (coef[m] - prev[m]);
            if (ghost_atom >= 0) ghost_val = ghost_prev - ss * ghost_prev;
          }
          done = true;
        } else if (n_iter >= P.max_iter || n_act >= k) {
          if (n_iter >= P.max_iter) status |= 4;
          done = true;
        } else {
          do_add = !drop;
        }
      }

      // ---- 2. atom j joins: border M ----
      if (__any_sync(0xffffffffu, do_add)) {
        const int j = bi;
        int cand = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          if (do_add && p < hw && acts[p] < 0 && p < cand) cand = p;
        }
        cand = gmini<LPC>(cand);
        const int qn = (cand == 0x7fffffff) ? hw : cand;
        if (do_add && qn >= SMAX) {       // active set outgrew this variant: hand the column to the large path
          status |= 8;
          done = true;
          do_add = false;
        }
        T gj[SA], u[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          T gv = T(0);
          if (do_add && p < hw) {
            int a = acts[p];
            if (a >= 0) gv = Gat(a, j);
          }
          gj[m] = gv;
          gs[p] = gv;
          u[m] = T(0);
        }
        __syncwarp();
        const int hwW = __reduce_max_sync(0xffffffffu, do_add ? hw : 0);
        for (int q = 0; q < hwW; ++q) {
          const T gq = gs[q];
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (do_add && q < hw && p < hw) u[m] += Mg[q * SMAX + p] * gq;
          }
        }
        T part = T(0);
#pragma unroll
        for (int m = 0; m < SA; ++m) part += gj[m] * u[m];
        const T Gjj = do_add ? Gat(j, j) : T(1);
        const T sig = Gjj - gsum<T, LPC>(part);
        T piv = sqrt(fabs(sig));
        piv = piv > piv_floor ? piv : piv_floor;
        bool degen = piv < T(1e-7);
        if (REFINE) degen = degen || !(sig > T(4) * eps32 * Gjj);   // fp32: Schur complement below rounding noise
        if (do_add && degen) {
          // degenerate regressor (sklearn _least_angle.py:723-742): covariance zeroed, atom stays inactive
          status |= 1;
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (l + LPC * m == j) cov[m] = T(0);
          do_add = false;
          skip = true;
        }
        const T inv = T(1) / (piv * piv);
#pragma unroll
        for (int m = 0; m < SA; ++m) us[l + LPC * m] = u[m];
        __syncwarp();
        for (int q = 0; q < hwW; ++q) {
          const T uq = us[q] * inv;
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (do_add && q < hw && p < hw) Mg[q * SMAX + p] += uq * u[m];
          }
        }
        __syncwarp();
        if (do_add) {
          const int hw_new = hw > qn + 1 ? hw : qn + 1;
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (p < hw_new) {
              T val = (p == qn) ? inv : -u[m] * inv;
              Mg[qn * SMAX + p] = val;
              Mg[p * SMAX + qn] = val;
              if (p == qn) { coef[m] = T(0); prev[m] = T(0); acts[qn] = j; }
            }
          }
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (l + LPC * m == j) inact &= ~(1u << m);
          hw = hw_new;
          ++n_act;
          max_act = n_act > max_act ? n_act : max_act;
        }
        __syncwarp();
      }

      // ---- 3. "alpha increasing" bail-out (sklearn _least_angle.py:752-765) ----
      if (!done && !skip && n_iter > 0 && a_prev < a_cur) {
        status |= 2;
        done = true;
      }
      bool live = !done && !skip;

      // ---- 4. equiangular weights w = AA * M 1 ----
      const int hwL = __reduce_max_sync(0xffffffffu, live ? hw : 0);
      T w[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) w[m] = T(0);
      for (int q = 0; q < hwL; ++q) {
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          if (live && q < hw && p < hw) w[m] += Mg[q * SMAX + p];
        }
      }
      if (REFINE) {
        // one step of iterative refinement: w += M (1 - G_AA w)
#pragma unroll
        for (int m = 0; m < SA; ++m) ws[l + LPC * m] = w[m];
        __syncwarp();
        T r[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          T rr = T(0);
          if (live && p < hw) {
            int ap = acts[p];
            if (ap >= 0) {
              T acc = T(0);
              for (int q = 0; q < hw; ++q) {
                int aq = acts[q];
                if (aq >= 0) acc += Gat(aq, ap) * ws[q];
              }
              rr = T(1) - acc;
            }
          }
          r[m] = rr;
        }
#pragma unroll
        for (int m = 0; m < SA; ++m) gs[l + LPC * m] = r[m];
        __syncwarp();
        for (int q = 0; q < hwL; ++q) {
          const T rq = gs[q];
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (live && q < hw && p < hw) w[m] += Mg[q * SMAX + p] * rq;
          }
        }
        __syncwarp();
      }
      T sw = T(0);
#pragma unroll
      for (int m = 0; m < SA; ++m) sw += w[m];
      sw = gsum<T, LPC>(sw);
      if (live && !(sw > T(0) && sw < Num<T>::big())) {   // active Gram block numerically singular
        status |= 16;
        done = true;
        live = false;
      }
      const T AA = live ? T(1) / sqrt(sw) : T(1);
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        w[m] *= AA;
        ws[l + LPC * m] = live ? w[m] : T(0);
      }
      __syncwarp();

      // ---- 5. correlation of every atom with the equiangular direction: G[:, A] w ----
      T corr[NA];
#pragma unroll
      for (int m = 0; m < NA; ++m) corr[m] = T(0);
      for (int q = 0; q < hwL; ++q) {
        const int a = (live && q < hw) ? acts[q] : -1;
        if (a >= 0) {
          const T wq = ws[q];
#pragma unroll
          for (int m = 0; m < NA; ++m) corr[m] += Gat(a, l + LPC * m) * wq;
        }
      }
      if (sizeof(T) == 8) {
        // np.around(corr_eq_dir, decimals=15)  (sklearn _least_angle.py:806)
#pragma unroll
        for (int m = 0; m < NA; ++m) corr[m] = T(rint(double(corr[m]) * 1e15) / 1e15);
      }

      // ---- 6. step length ----
      T g1 = Num<T>::big();
#pragma unroll
      for (int m = 0; m < NA; ++m)
        if ((inact >> m) & 1u) {
          T v = (C - cov[m]) / (AA - corr[m] + tiny);
          if (v > T(0) && v < g1) g1 = v;
        }
      g1 = gmin<T, LPC>(g1);
      T gamma = C / AA;
      gamma = g1 < gamma ? g1 : gamma;
      T zbest = Num<T>::big();
      int zs = 0x7fffffff;
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        int p = l + LPC * m;
        if (live && p < hw && acts[p] >= 0) {
          T z = -coef[m] / (w[m] + tiny);
          if (z > T(0) && z < zbest) { zbest = z; zs = p; }
        }
      }
#pragma unroll
      for (int off = LPC / 2; off > 0; off >>= 1) {
        T ov = __shfl_xor_sync(0xffffffffu, zbest, off);
        int oi = __shfl_xor_sync(0xffffffffu, zs, off);
        if (ov < zbest || (ov == zbest && oi > zs && oi != 0x7fffffff)) { zbest = ov; zs = oi; }
      }
      if (live) {
        drop = false;
        if (zbest < gamma) { gamma = zbest; drop = true; dslot = zs; }
        // ---- 7. move along the path ----
        ++n_iter;
        a_prev = a_cur;
        ghost_atom = -1;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          prev[m] = coef[m];
          coef[m] = prev[m] + gamma * w[m];
        }
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if ((inact >> m) & 1u) cov[m] -= gamma * corr[m];
        ++st_knots;
        st_s += (unsigned)n_act;
        st_s2 += (unsigned)(n_act * n_act);
      }

      // ---- 8. atom leaves: Schur downdate of M, exact covariance of the dropped atom ----
      const bool dodrop = live && drop;
      if (__any_sync(0xffffffffu, dodrop)) {
        const int p0 = dodrop ? dslot : 0;
        const int a_d = dodrop ? acts[p0] : 0;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          us[p] = (dodrop && p < hw) ? Mg[p0 * SMAX + p] : T(0);
        }
        __syncwarp();
        const T mpp = dodrop ? us[p0] : T(1);
        {
          T gp = T(0);
#pragma unroll
          for (int m = 0; m < SA; ++m)
            if (dodrop && l + LPC * m == p0) gp = prev[m];
          gp = gsum<T, LPC>(gp);
          if (dodrop) { ghost_atom = a_d; ghost_prev = gp; }
        }
        const int hwD = __reduce_max_sync(0xffffffffu, dodrop ? hw : 0);
        T ur[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) ur[m] = us[l + LPC * m];
        for (int q = 0; q < hwD; ++q) {
          const T f = us[q] / mpp;
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (dodrop && q < hw && p < hw) Mg[q * SMAX + p] -= f * ur[m];
          }
        }
        __syncwarp();
        if (dodrop) {
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            int p = l + LPC * m;
            if (p < hw) {
              Mg[p0 * SMAX + p] = T(0);
              Mg[p * SMAX + p0] = T(0);
            }
            if (p == p0) { coef[m] = T(0); prev[m] = T(0); acts[p0] = -1; }
          }
          --n_act;
          ++st_drops;
        }
        __syncwarp();
        int top = 0;
        T part = T(0);
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          int p = l + LPC * m;
          if (dodrop && p < hw) {
            int a = acts[p];
            if (a >= 0) {
              top = p + 1;
              part += Gat(a_d, a) * coef[m];
            }
          }
        }
        top = gmaxi<LPC>(top);
        part = gsum<T, LPC>(part);
        if (dodrop) {
          hw = top;
#pragma unroll
          for (int m = 0; m < NA; ++m)
            if (l + LPC * m == a_d) {
              cov[m] = crow[a_d] - part;
              inact |= 1u << m;
            }
        }
      }
    }  // path loop

    // ---- write the code row ----
    const bool ovf = (status & 8) != 0;
    if (valid && !ovf) {
      T* hrow = P.Ht + (size_t)col * k;
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        int i = l + LPC * m;
        if (i < k) hrow[i] = T(0);
      }
    }
    __syncwarp();
    if (valid && !ovf) {
      T* hrow = P.Ht + (size_t)col * k;
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        int p = l + LPC * m;
        if (p < hw) {
          int a = acts[p];
          if (a >= 0) hrow[a] = coef[m];
        }
      }
      if (l == 0 && ghost_atom >= 0 && ghost_val != T(0)) hrow[ghost_atom] = ghost_val;
    }
    if (valid && l == 0) {
      if (ovf) {
        if (P.ovf_list) {
          unsigned slot = atomicAdd(P.ovf_count, 1u);
          P.ovf_list[slot] = col;
        }
        ++st_ovf;
      } else {
        ++st_cols;
        if (status & ~8) ++st_flag;
      }
      st_maxact = max_act > st_maxact ? max_act : st_maxact;
    }
    __syncwarp();
  }  // ticket loop

  if (P.stats && l == 0) {
    // lanes other than the group leader carry zero counters except knots/s/s2 (group-uniform): leader only
    atomicAdd(&P.stats->columns, st_cols);
    atomicAdd(&P.stats->knots, st_knots);
    atomicAdd(&P.stats->sum_active, st_s);
    atomicAdd(&P.stats->sum_active2, st_s2);
    atomicAdd(&P.stats->drops, st_drops);
    atomicAdd(&P.stats->overflow, st_ovf);
    atomicAdd(&P.stats->flagged, st_flag);
    atomicMax(&P.stats->max_active, (unsigned long long)st_maxact);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

struct LarsWs {           // header at the start of the caller's workspace
  unsigned long long ticket_main;
  unsigned long long ticket_ovf;
  unsigned int ovf_count;
  unsigned int pad[11];
};
static_assert(sizeof(LarsWs) == 64, "header size");

static int k_class(int k) { return k <= 32 ? 0 : k <= 64 ? 1 : k <= 128 ? 2 : k <= 256 ? 3 : k <= 512 ? 4 : -1; }
static int class_kp(int c) { static const int kp[5] = {32, 64, 128, 256, 512}; return kp[c]; }
static int class_lpc(int c) { static const int v[5] = {8, 16, 32, 32, 32}; return v[c]; }

static size_t ovf_scratch_groups(int kp, size_t tsz) {
  size_t per = (size_t)kp * kp * tsz;
  size_t g = (96ull << 20) / per;
  if (g < 16) g = 16;
  if (g > 592) g = 592;
  return g;
}

template <typename T, int LPC, int NA, int SMAX>
static int launch_class(const T* G, const T* Ct, long long n, int k, int d, double alpha, int max_iter, T* Ht,
                        unsigned char* ws, onmf_lars_stats* stats, cudaStream_t st) {
  constexpr int GPW = 32 / LPC;
  constexpr int KP = LPC * NA;
  LarsWs* hdr = reinterpret_cast<LarsWs*>(ws);
  long long* ovf_list = reinterpret_cast<long long*>(ws + sizeof(LarsWs));
  size_t list_bytes = round_up<size_t>((size_t)n * sizeof(long long), 256);
  T* mscr = reinterpret_cast<T*>(ws + sizeof(LarsWs) + list_bytes);
  ONMF_CUDA(cudaMemsetAsync(hdr, 0, sizeof(LarsWs), st));

  LarsParams<T> P;
  P.G = G; P.Ct = Ct; P.Ht = Ht; P.n = n; P.k = k; P.d = d; P.max_iter = max_iter;
  P.amin = T(alpha) / T(d);
  P.ticket = &hdr->ticket_main; P.col_list = nullptr; P.n_list = nullptr;
  P.ovf_list = ovf_list; P.ovf_count = &hdr->ovf_count; P.Mscratch = nullptr; P.stats = stats;

  const int smem_max = max_smem_optin();
  const size_t g_bytes = round_up<size_t>((size_t)k * gram_stride<LPC, NA>() * sizeof(T), 128);
  const size_t grp_bytes = (size_t)group_words<T, LPC, SMAX, false>() * 4;
  const bool gsm = g_bytes <= 72 * 1024 && (smem_max - (long)g_bytes) >= (long)(2 * GPW * grp_bytes);
  const size_t avail = smem_max - (gsm ? g_bytes : 0) - 256;
  int nw = (int)(avail / (GPW * grp_bytes));
  if (nw > 16) nw = 16;
  if (nw < 1) return fail(ONMF_E_UNSUPPORTED, "lasso_lars: shared memory too small for one warp");
  // spread small minibatches over all SMs instead of packing few CTAs
  long long groups = cdiv<long long>(n, GPW);
  int nw_need = (int)cdiv<long long>(groups, num_sms());
  if (nw_need < nw) nw = nw_need < 1 ? 1 : nw_need;
  int grid = (int)cdiv<long long>(groups, nw);
  if (grid > num_sms()) grid = num_sms();
  size_t smem = (gsm ? g_bytes : 0) + (size_t)nw * GPW * grp_bytes;
  if (gsm) {
    auto kern = lars_kernel<T, LPC, NA, SMAX, true, false>;
    ONMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, nw * 32, smem, st>>>(P);
  } else {
    auto kern = lars_kernel<T, LPC, NA, SMAX, false, false>;
    ONMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, nw * 32, smem, st>>>(P);
  }
  ONMF_LAUNCH_CHECK("lars_kernel");

  if (SMAX < KP) {
    // large-active-set pass over the columns the main pass could not finish (device-side list)
    LarsParams<T> Q = P;
    Q.ticket = &hdr->ticket_ovf; Q.col_list = ovf_list; Q.n_list = &hdr->ovf_count;
    Q.ovf_list = nullptr; Q.ovf_count = nullptr; Q.Mscratch = mscr;
    size_t ng = ovf_scratch_groups(KP, sizeof(T));
    const int nw2 = 2;
    int grid2 = (int)(ng / (nw2 * GPW));
    if (grid2 < 1) grid2 = 1;
    long long need = cdiv<long long>(n, nw2 * GPW);
    if (grid2 > need) grid2 = (int)need;
    size_t smem2 = (size_t)nw2 * GPW * group_words<T, LPC, KP, true>() * 4;
    auto kern2 = lars_kernel<T, LPC, NA, KP, false, true>;
    ONMF_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    kern2<<<grid2, nw2 * 32, smem2, st>>>(Q);
    ONMF_LAUNCH_CHECK("lars_kernel(overflow)");
  }
  return ONMF_OK;
}

template <typename T>
static int lasso_lars_t(const void* G, const void* Ct, long long n, int k, int d, double alpha, int max_iter,
                        void* Ht, void* ws, onmf_lars_stats* stats, cudaStream_t st) {
  const T* g = (const T*)G; const T* c = (const T*)Ct; T* h = (T*)Ht; unsigned char* w = (unsigned char*)ws;
  switch (k_class(k)) {
    case 0: return launch_class<T, 8, 4, 32>(g, c, n, k, d, alpha, max_iter, h, w, stats, st);
    case 1: return launch_class<T, 16, 4, 64>(g, c, n, k, d, alpha, max_iter, h, w, stats, st);
    case 2: return launch_class<T, 32, 4, 64>(g, c, n, k, d, alpha, max_iter, h, w, stats, st);
    case 3: return launch_class<T, 32, 8, 64>(g, c, n, k, d, alpha, max_iter, h, w, stats, st);
    case 4: return launch_class<T, 32, 16, 64>(g, c, n, k, d, alpha, max_iter, h, w, stats, st);
  }
  return fail(ONMF_E_UNSUPPORTED, "lasso_lars: n_components > 512 not instantiated");
}

}  // namespace onmf

extern "C" size_t onmf_lasso_lars_workspace(int dtype, int k, int64_t n) {
  int c = onmf::k_class(k);
  if (c < 0 || n < 0) return 0;
  size_t tsz = dtype == ONMF_F64 ? 8 : 4;
  int kp = onmf::class_kp(c);
  size_t bytes = sizeof(onmf::LarsWs) + onmf::round_up<size_t>((size_t)n * sizeof(long long), 256);
  int smax_main = c == 0 ? 32 : 64;
  if (smax_main < kp) bytes += onmf::ovf_scratch_groups(kp, tsz) * (size_t)kp * kp * tsz;
  return bytes + 256;
}

extern "C" int onmf_lasso_lars(int dtype, const void* G, const void* Ct, int64_t n, int k, int d, double alpha,
                               int max_iter, void* Ht, void* workspace, size_t workspace_bytes,
                               onmf_lars_stats* stats, void* stream) {
  using namespace onmf;
  if (!G || !Ct || !Ht || !workspace) return fail(ONMF_E_ARG, "lasso_lars: null pointer");
  if (n < 0 || k <= 0 || d <= 0 || max_iter < 0 || !(alpha >= 0.0)) return fail(ONMF_E_ARG, "lasso_lars: bad size/alpha");
  if (dtype != ONMF_F32 && dtype != ONMF_F64) return fail(ONMF_E_ARG, "lasso_lars: bad dtype");
  if (n == 0) return ONMF_OK;
  if (workspace_bytes < onmf_lasso_lars_workspace(dtype, k, n)) return fail(ONMF_E_WORKSPACE, "lasso_lars: workspace too small");
  if ((uintptr_t)workspace % 256) return fail(ONMF_E_ARG, "lasso_lars: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == ONMF_F32) return lasso_lars_t<float>(G, Ct, n, k, d, alpha, max_iter, Ht, workspace, stats, st);
  return lasso_lars_t<double>(G, Ct, n, k, d, alpha, max_iter, Ht, workspace, stats, st);
}
