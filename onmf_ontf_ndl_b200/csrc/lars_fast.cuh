// K3, fast first tier -- included by lars.cu (inside namespace onmf).
//
// The fp32 production coder for k > 128 (one warp per column, 64 slots, hybrid shared/global factor) rewritten around
// WARP-UNIFORM control flow: with one column per warp every path decision (join / drop / stop) is the same for all 32
// lanes, so the knot loop is plain branches instead of the predicated, vote-guarded form the general kernel
// (lars_kernel: several columns per warp, both precisions, every k class) needs.  It walks the CLEAN homotopy path only:
// a column that meets any of sklearn's special events -- degenerate pivot (_least_angle.py:723-742), "alpha increasing"
// bail-out (:752-765), max_iter, a numerically singular active block -- or outgrows the 64 slots is handed, untouched, to
// the next tier (the general kernel, which re-walks it from the start with the full semantics); a few columns in 10^5.
// The arithmetic of a clean path is the general kernel's, operation for operation (same reduction trees, same
// accumulation order), so both produce the same bits; tests/test_gpu_parity.py::test_fast_tier_matches_general_kernel.
//
// Per knot (s = active atoms, NA = k/32 atoms per lane):
//   arg-max of the inactive covariances   NA FMNMX + 2 REDUX
//   join: g = G64[j, A] (one gather), t = V^T g, u = V t over the packed FP64 inverse factor (shared memory, columns
//         >= SPLIT in an L2-resident tail), |t|^2 and 1^T u reduced TOGETHER (one butterfly with two values in flight)
//   equiangular weights (incremental), normalisation, slot table
//   correlation pass G[:, A] w: float4 Gram-row loads UQ rows deep, packed FFMA2 (two fp32 FMAs per instruction)
//   step length: packed add / mul, MUFU.RCP, REDUX min
//   a drop (one per ~30 knots) runs the general kernel's Givens downdate

// a warp-uniform decision, stated so that the compiler can see it (a vote result is uniform by construction): keeps the
// *_sync collectives below it out of the divergent-code trampolines
#define UNI(cond) __any_sync(0xffffffffu, (cond))

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int NA, int SMAX, int SPLIT>
__global__ void __launch_bounds__(LARS_MAX_THREADS, 1) lars_fast_kernel(LarsParams<float> P) {
  typedef float T;
  constexpr int SA = SMAX / 32;          // slot registers per lane (slot p = l + 32 m)
  constexpr int NV = NA / 4;             // float4 loads per Gram row per lane
  constexpr int KP = 32 * NA;            // padded row length of the Gram copy
  constexpr int UQ = (NA >= 16) ? 2 : 4; // Gram rows in flight per batch of the correlation pass
  static_assert(SMAX % 32 == 0 && NA % 4 == 0 && SPLIT >= 32 && SPLIT < SMAX, "bad tile shape");

  if (P.hint != nullptr && *P.hint != (unsigned)P.run_if) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k;
  const int l = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint32_t* gbase = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * group_words<T, 32, SMAX, false, SPLIT>();
  constexpr int MELEMS = SMAX * (SMAX + 1) / 2;
  constexpr int MSM = SPLIT * (SPLIT + 1) / 2;
  double* Mg = reinterpret_cast<double*>(gbase);                                   // packed columns 0..SPLIT-1 of V
  double* Mx = P.Mhyb + ((size_t)blockIdx.x * (blockDim.x / 32) + warp) * (size_t)(MELEMS - MSM);   // columns >= SPLIT
  double* gs = Mg + MSM;                                                           // g = G[A, j]
  double* us = gs + SMAX;                                                          // t = V^T g
  SlotW<T>* sw_ = reinterpret_cast<SlotW<T>*>(us + SMAX);                          // (atom, weight) by slot
  int* acts = reinterpret_cast<int*>(sw_ + SMAX);                                  // slot -> atom

  auto atom_of = [&](int m) -> int { return ((m >> 2) * 32 + l) * 4 + (m & 3); };
  unsigned long long Grl = reinterpret_cast<unsigned long long>(P.Gp) + (unsigned long long)l * sizeof(float4);
  asm volatile("" : "+l"(Grl));
  const double* __restrict__ G64 = P.G64;

  const T tiny = T(1.1754943508222875e-38);
  const T dT = T(P.d);
  const T eps32 = T(1.1920928955078125e-07) * dT;     // covariance units (see lars_kernel)
  const T amin = P.amin * dT;
  const T NINF = -CUDART_INF_F;
  const T BIG = 3.402823466e+38f;

  int ci[SA];
#pragma unroll
  for (int m = 0; m < SA; ++m) { const int i = l + 32 * m; ci[m] = i * (i + 1) / 2; }
  auto Vld = [&](int idx) -> double { return (idx >= MSM) ? Mx[idx - MSM] : Mg[idx]; };
  auto Vst = [&](int idx, double v) { if (idx >= MSM) Mx[idx - MSM] = v; else Mg[idx] = v; };

  // t_i = sum_{p <= i} V[p][i] src[p], i < s (this lane's slots)
  auto sweep_t = [&](double (&t)[SA], const double* src, int s) {
    if (s <= 32) {
      const double* col = Mg + ci[0];
      const int ie = (l < s) ? l : -1;
#pragma unroll 4
      for (int p = 0; p < s; ++p) {
        const double sp = src[p];
        if (p <= ie) t[0] += col[p] * sp;
      }
    } else if (s <= SPLIT) {
      int ie[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) { const int i = l + 32 * m; ie[m] = (i < s) ? i : -1; }
#pragma unroll 4
      for (int p = 0; p < s; ++p) {
        const double sp = src[p];
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (p <= ie[m]) t[m] += Mg[ci[m] + p] * sp;
      }
    } else {
      int ie[SA];
      const double* colp[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        const int i = l + 32 * m;
        ie[m] = (i < s) ? i : -1;
        colp[m] = (i >= SPLIT) ? (Mx + (ci[m] - MSM)) : (Mg + ci[m]);
      }
#pragma unroll 2
      for (int p = 0; p < s; ++p) {
        const double sp = src[p];
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (p <= ie[m]) t[m] += colp[m][p] * sp;
      }
    }
  };
  // u_p = sum_{p <= i < s} V[p][i] src[i]
  auto sweep_u = [&](double (&u)[SA], const double* src, int s) {
    const int nS = s > SPLIT ? SPLIT : s;
    const double* rp = Mg + l;
    const int nS1 = nS < 32 ? nS : 32;
#pragma unroll 4
    for (int i = 0; i < nS1; ++i) {
      const double si = src[i];
      if (l <= i) u[0] += rp[0] * si;
      rp += i + 1;
    }
#pragma unroll 2
    for (int i = nS1; i < nS; ++i) {
      const double si = src[i];
#pragma unroll
      for (int m = 0; m < SA; ++m)
        if (l + 32 * m <= i) u[m] += rp[32 * m] * si;
      rp += i + 1;
    }
    if (s > SPLIT) {
      rp = Mx + l;
#pragma unroll 2
      for (int i = SPLIT; i < s; ++i) {
        const double si = src[i];
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (l + 32 * m <= i) u[m] += rp[32 * m] * si;
        rp += i + 1;
      }
    }
  };

  unsigned long long st_knots = 0, st_s = 0, st_s2 = 0, st_drops = 0, st_cols = 0, st_ovf = 0;
  int st_maxact = 0;

  while (true) {
    unsigned long long widx = 0;
    if (l == 0) widx = atomicAdd(P.ticket, 1ull);
    widx = __shfl_sync(0xffffffffu, widx, 0);
    const unsigned long long nwork = P.col_list ? (unsigned long long)(*P.n_list) : (unsigned long long)P.n;
    if (UNI(widx >= nwork)) break;
    const long long col = P.col_list ? P.col_list[widx] : (long long)widx;
    const T* crow = P.Ct + (size_t)col * k;

    // active atoms and the padding beyond k carry cov = -inf (see lars_kernel)
    T cov[NA];
    if (k == KP) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 c4 = *reinterpret_cast<const float4*>(crow + (v * 32 + l) * 4);
        cov[4 * v + 0] = c4.x; cov[4 * v + 1] = c4.y; cov[4 * v + 2] = c4.z; cov[4 * v + 3] = c4.w;
      }
    } else {
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const int i = atom_of(m);
        cov[m] = (i < k) ? crow[i] : NINF;
      }
    }
    T coef[SA], prev[SA];
    double wd[SA];
#pragma unroll
    for (int m = 0; m < SA; ++m) {
      coef[m] = T(0); prev[m] = T(0); wd[m] = 0.0;
      SlotW<T> z; z.atom = 0; z.w = T(0);
      sw_[l + 32 * m] = z;
      acts[l + 32 * m] = -1;
    }
    double sw = 0.0;
    int n_iter = 0, n_act = 0, max_act = 0;
    unsigned kn_s = 0, kn_s2 = 0;
    bool drop = false;
    bool handoff = false;                // column leaves this tier unfinished
    bool slots_full = false;
    int dslot = 0;
    T a_prev = T(0);
    int ghost_atom = -1;
    T ghost_prev = T(0), ghost_val = T(0);
    int banned = -1;
    __syncwarp();

    while (true) {
      // ---- 1. largest inactive covariance (value first, then the lowest atom attaining it) ----
      T best = NINF;
#pragma unroll
      for (int m = 0; m < NA; ++m) best = cov[m] > best ? cov[m] : best;
      best = gmaxval<32>(best, 0xffffffffu);
      int bi = 0x7fffffff;
#pragma unroll
      for (int m = NA - 1; m >= 0; --m)
        if (cov[m] == best) bi = atom_of(m);
      bi = __reduce_min_sync(0xffffffffu, bi);
      if (UNI(banned >= 0)) {
        // the atom dropped by the last drop step cannot be the joiner of the first join knot after it (see lars_kernel)
        if (UNI(!drop && bi == banned)) {
          T b2 = NINF;
          int i2 = 0x7fffffff;
#pragma unroll
          for (int m = 0; m < NA; ++m) {
            const T cv = (atom_of(m) == banned) ? NINF : cov[m];
            b2 = cv > b2 ? cv : b2;
          }
          b2 = gmaxval<32>(b2, 0xffffffffu);
#pragma unroll
          for (int m = NA - 1; m >= 0; --m)
            if (cov[m] == b2 && atom_of(m) != banned) i2 = atom_of(m);
          i2 = __reduce_min_sync(0xffffffffu, i2);
          if (b2 > NINF) { best = b2; bi = i2; }
        }
        if (!drop) banned = -1;
      }
      const T C = best > NINF ? best : T(0);
      // ---- stopping rule: alpha reached -> interpolate inside the last segment ----
      if (UNI(C <= amin + eps32)) {
        const T diff = C - amin;
        if ((diff > eps32 || diff < -eps32) && n_iter > 0) {
          const T ss = (a_prev - amin) / (a_prev - C);
#pragma unroll
          for (int m = 0; m < SA; ++m) coef[m] = prev[m] + ss * (coef[m] - prev[m]);
          if (ghost_atom >= 0) ghost_val = ghost_prev - ss * ghost_prev;
        }
        break;
      }
      if (UNI(n_iter >= P.max_iter || n_act >= k)) { handoff = true; break; }

      // ---- 2. atom j joins slot n_act ----
      if (UNI(!drop)) {
        if (UNI(n_act >= SMAX)) { handoff = true; slots_full = true; break; }
        const int j = bi;
        const int s = n_act;
        const double gjj = G64[(size_t)j * k + j];
        double t[SA], u[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          t[m] = 0.0; u[m] = 0.0;
          if (32 * m < s) {
            const int p = l + 32 * m;
            double gv = 0.0;
            if (p < s) gv = G64[(size_t)j * k + acts[p]];
            gs[p] = gv;
          }
        }
        __syncwarp();
        sweep_t(t, gs, s);
        double tt = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          tt += t[m] * t[m];
          if (32 * m < s) us[l + 32 * m] = t[m];
        }
        __syncwarp();
        sweep_u(u, us, s);
        double su = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) su += u[m];
        // |t|^2 and 1^T u: one butterfly, two values in flight
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double a = __shfl_xor_sync(0xffffffffu, tt, off);
          const double b = __shfl_xor_sync(0xffffffffu, su, off);
          tt += a; su += b;
        }
        const double sig = gjj - tt;
        double asig = fabs(sig);
        asig = asig > 4.930380657631324e-32 ? asig : 4.930380657631324e-32;
        if (UNI(asig < 1e-14)) { handoff = true; break; }          // degenerate regressor: the general kernel handles it
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (atom_of(m) == j) cov[m] = NINF;
        const double rs = fast_rsqrt(asig);
        const double tau = (1.0 - su) * rs * rs;
        const int cs = s * (s + 1) / 2;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          if (p <= s) {
            Vst(cs + p, (p == s) ? rs : -u[m] * rs);
            wd[m] = (p == s) ? tau : wd[m] - tau * u[m];
            if (p == s) acts[p] = j;
          }
        }
        sw += tau * (1.0 - su);
        n_act = s + 1;
        max_act = n_act > max_act ? n_act : max_act;
        __syncwarp();
      }

      // the Gram rows of the first batch belong to slots 0..UQ-1, final once the join is done: request them now
      float4 gv0[UQ][NV];
#pragma unroll
      for (int t = 0; t < UQ; ++t) {
        const int a0 = acts[t];
        const float4* row = reinterpret_cast<const float4*>(Grl + (unsigned long long)(unsigned)(a0 >= 0 ? a0 : 0) * (unsigned)(KP * sizeof(T)));
#pragma unroll
        for (int v = 0; v < NV; ++v) gv0[t][v] = ld_global_vec(row + v * 32);
      }
      // ---- 3. "alpha increasing" (sklearn _least_angle.py:752-765), singular block: general kernel ----
      if (UNI((n_iter > 0 && a_prev < C) || !(sw > 0.0 && sw < 1e300))) { handoff = true; break; }

      // ---- 4. normalise the equiangular weights ----
      const double AAd = fast_rsqrt1(sw);
      const T AA = (T)AAd;
      T w[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        w[m] = (T)(wd[m] * AAd);
        if (32 * m < n_act) {
          const int p = l + 32 * m;
          if (p < n_act) {
            SlotW<T> e; e.atom = acts[p]; e.w = w[m];
            sw_[p] = e;
          }
        }
      }
      __syncwarp();

      // ---- 5. correlation of every atom with the equiangular direction ----
      float2 corr2[NA / 2];
#pragma unroll
      for (int m = 0; m < NA / 2; ++m) corr2[m] = make_float2(0.f, 0.f);
      {
        SlotW<T> e[UQ];
#pragma unroll
        for (int t = 0; t < UQ; ++t) e[t] = sw_[t];
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
          const float2 ww = make_float2(e[t].w, e[t].w);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            corr2[2 * v] = ffma2(make_float2(gv0[t][v].x, gv0[t][v].y), ww, corr2[2 * v]);
            corr2[2 * v + 1] = ffma2(make_float2(gv0[t][v].z, gv0[t][v].w), ww, corr2[2 * v + 1]);
          }
        }
      }
      for (int q0 = UQ; q0 < n_act; q0 += UQ) {
        SlotW<T> e[UQ];
#pragma unroll
        for (int t = 0; t < UQ; ++t) e[t] = sw_[q0 + t];
        float4 gv[UQ][NV];
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
          const float4* row = reinterpret_cast<const float4*>(Grl + (unsigned long long)(unsigned)e[t].atom * (unsigned)(KP * sizeof(T)));
#pragma unroll
          for (int v = 0; v < NV; ++v) gv[t][v] = ld_global_vec(row + v * 32);
        }
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
          const float2 ww = make_float2(e[t].w, e[t].w);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            corr2[2 * v] = ffma2(make_float2(gv[t][v].x, gv[t][v].y), ww, corr2[2 * v]);
            corr2[2 * v + 1] = ffma2(make_float2(gv[t][v].z, gv[t][v].w), ww, corr2[2 * v + 1]);
          }
        }
      }
      T corr[NA];
#pragma unroll
      for (int m = 0; m < NA / 2; ++m) { corr[2 * m] = corr2[m].x; corr[2 * m + 1] = corr2[m].y; }

      // ---- 6. step length ----
      T g1 = BIG;
#pragma unroll
      for (int m = 0; m < NA; ++m) {
        const T den = AA - corr[m] + tiny;
        const T num = C - cov[m];
        const T v = qdiv(num > T(0) ? num : T(0), den);
        if (den > T(0) && v < g1) g1 = v;
      }
      g1 = gminpos<32>(g1, 0xffffffffu);
      T gamma = qdiv(C, AA);
      gamma = g1 < gamma ? g1 : gamma;
      T zbest = BIG;
      int zs = -1;
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        if (32 * m < n_act) {
          const int p = l + 32 * m;
          if (p < n_act) {
            const T z = qdiv(-coef[m], w[m] + tiny);
            if (z > T(0) && z < zbest) { zbest = z; zs = p; }
          }
        }
      }
      gargminpos<32>(zbest, zs, 0xffffffffu);
      drop = UNI(zbest < gamma && zs >= 0);
      if (drop) { gamma = zbest; dslot = zs; }
      // ---- 7. move along the path ----
      ++n_iter;
      a_prev = C;
      ghost_atom = -1;
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        prev[m] = coef[m];
        coef[m] = prev[m] + gamma * w[m];
      }
#pragma unroll
      for (int m = 0; m < NA; ++m) cov[m] -= gamma * corr[m];
      kn_s += (unsigned)n_act;
      kn_s2 += (unsigned)(n_act * n_act);

      // ---- 8. atom leaves slot p0 (Givens downdate of the factor; see lars_kernel) ----
      if (drop) {
        const int p0 = dslot;
        const int a_d = acts[dslot];
        const int sD = n_act;
        double rr[SA], mv[SA], pre[SA];
        T gp = T(0);
        double wp0 = 0.0, rp0 = 0.0, mpp = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int i = l + 32 * m;
          rr[m] = (i >= p0 && i < n_act) ? Vld(ci[m] + p0) : 0.0;
          mv[m] = 0.0;
          if (32 * m < sD) us[i] = rr[m];
          if (i == p0) { gp = prev[m]; wp0 = wd[m]; rp0 = rr[m]; }
          mpp += rr[m] * rr[m];
        }
        gp = gsum<32>(gp);
        wp0 = gsum<32>(wp0);
        rp0 = gsum<32>(rp0);
        mpp = gsum<32>(mpp);
        ghost_atom = a_d; ghost_prev = gp; banned = a_d;
        __syncwarp();
        sweep_u(mv, us, sD);
        double sum_m = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) sum_m += mv[m];
        sum_m = gsum<32>(sum_m);
        const double fw = wp0 * fast_rcp(mpp);
        {
          double run = 0.0;
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            double x = rr[m] * rr[m];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const double y = __shfl_up_sync(0xffffffffu, x, off, 32);
              if (l >= off) x += y;
            }
            pre[m] = x + run;
            run += __shfl_sync(0xffffffffu, x, 31, 32);
          }
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int i1 = l + 32 * m;
          if (32 * m < sD) {
            double c = 1.0, sn = 0.0;
            if (i1 > p0 && i1 < n_act) {
              const double rinv = fast_rsqrt(pre[m]);
              const double prev2 = pre[m] - rr[m] * rr[m];
              const double a = (i1 - 1 == p0) ? rp0 : prev2 * fast_rsqrt(prev2);
              c = rr[m] * rinv;
              sn = -a * rinv;
            }
            gs[i1] = c;
            us[i1] = sn;
          }
        }
        __syncwarp();
        {
          double R[SA];
          int nrow[SA];
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            const int p = l + 32 * m;
            R[m] = (p <= p0) ? Vld(p0 * (p0 + 1) / 2 + p) : 0.0;
            nrow[m] = (p < p0) ? p : (p == p0 ? -1 : p - 1);
          }
          __syncwarp();
          for (int i = p0; i + 1 < sD; ++i) {
            const double c = gs[i + 1], sn = us[i + 1];
            const int cn = (i + 1) * (i + 2) / 2, co = i * (i + 1) / 2;
#pragma unroll
            for (int m = 0; m < SA; ++m) {
              if (32 * m <= i + 1) {
                const int p = l + 32 * m;
                if (p <= i + 1) {
                  const double X = Vld(cn + p);
                  const double nc = c * R[m] + sn * X;
                  R[m] = c * X - sn * R[m];
                  if (nrow[m] >= 0) Vst(co + nrow[m], nc);
                }
              }
            }
            __syncwarp();
          }
        }
        int nxt_act[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          nxt_act[m] = (p + 1 < SMAX) ? acts[p + 1] : -1;
          wd[m] = (p < n_act && p != p0) ? wd[m] - fw * mv[m] : 0.0;
        }
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          const T c_dn = __shfl_down_sync(0xffffffffu, coef[m], 1, 32);
          const T p_dn = __shfl_down_sync(0xffffffffu, prev[m], 1, 32);
          const double w_dn = __shfl_down_sync(0xffffffffu, wd[m], 1, 32);
          const T c_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? coef[m + 1 < SA ? m + 1 : m] : T(0), 0, 32);
          const T p_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? prev[m + 1 < SA ? m + 1 : m] : T(0), 0, 32);
          const double w_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? wd[m + 1 < SA ? m + 1 : m] : 0.0, 0, 32);
          if (p >= p0) {
            coef[m] = (l == 31) ? c_wr : c_dn;
            prev[m] = (l == 31) ? p_wr : p_dn;
            wd[m] = (l == 31) ? w_wr : w_dn;
          }
        }
        __syncwarp();
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          if (p >= p0) acts[p] = nxt_act[m];
        }
        sw = sw - wp0 - fw * (sum_m - mpp);
        --n_act;
        ++st_drops;
        if (l == 0) { SlotW<T> z; z.atom = 0; z.w = T(0); sw_[n_act] = z; }
        __syncwarp();
        // exact covariance of the dropped atom (sklearn _least_angle.py:891)
        T part = T(0);
        const T* grow = P.Gp + (size_t)a_d * KP;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          if (32 * m < n_act) {
            const int p = l + 32 * m;
            if (p < n_act) part += grow[acts[p]] * coef[m];
          }
        }
        part = gsum<32>(part);
        const T cnew = crow[a_d] - part;
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (atom_of(m) == a_d) cov[m] = cnew;
        __syncwarp();
      }
    }  // path loop

    // ---- write the code row / hand the column on ----
    if (UNI(!handoff)) {
      T* hrow = P.Ht + (size_t)col * k;
      if (k == KP) {
#pragma unroll
        for (int v = 0; v < NV; ++v) *reinterpret_cast<float4*>(hrow + (v * 32 + l) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int m = 0; m < NA; ++m) {
          const int i = atom_of(m);
          if (i < k) hrow[i] = T(0);
        }
      }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        const int p = l + 32 * m;
        if (p < n_act) hrow[acts[p]] = coef[m];
      }
      if (l == 0 && ghost_atom >= 0 && ghost_val != T(0)) hrow[ghost_atom] = ghost_val;
      if (l == 0) {
        ++st_cols;
        if (P.over_thresh != nullptr && max_act > P.thresh) atomicAdd(P.over_thresh, 1u);
      }
    } else if (l == 0) {
      if (P.ovf_list) {
        const unsigned slot = atomicAdd(P.ovf_count, 1u);
        P.ovf_list[slot] = col;
      }
      if (P.count_stats && slots_full) ++st_ovf;
    }
    st_maxact = max_act > st_maxact ? max_act : st_maxact;
    st_knots += (unsigned)n_iter;
    st_s += kn_s;
    st_s2 += kn_s2;
    __syncwarp();
  }  // ticket loop

  if (P.stats && l == 0) {
    atomicAdd(&P.stats->columns, st_cols);
    atomicAdd(&P.stats->knots, st_knots);
    atomicAdd(&P.stats->sum_active, st_s);
    atomicAdd(&P.stats->sum_active2, st_s2);
    atomicAdd(&P.stats->drops, st_drops);
    atomicAdd(&P.stats->overflow, st_ovf);
    atomicMax(&P.stats->max_active, (unsigned long long)st_maxact);
  }
}

#undef UNI
