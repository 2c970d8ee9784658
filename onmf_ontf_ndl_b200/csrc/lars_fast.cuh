// K3, fast first tier -- included by lars.cu (inside namespace onmf).
//
// The fp32 production coder for k > 64 (one warp per column) rewritten around WARP-UNIFORM control flow: with one column
// per warp every path decision (join / drop / stop) is the same for all 32 lanes, so the knot loop is plain branches
// instead of the predicated, vote-guarded form the general kernel (lars_kernel: several columns per warp, both precisions,
// every k class) needs.  It walks the CLEAN homotopy path only: a column that meets any of sklearn's special events --
// degenerate pivot (_least_angle.py:723-742), "alpha increasing" bail-out (:752-765), max_iter, a numerically singular active
// block -- or outgrows the slots of this tier is handed, untouched, to the next tier (the general kernel, which re-walks it
// from the start with the full semantics); a few columns in 10^5.  The arithmetic of a clean path is the general kernel's,
// operation for operation (same reduction trees, same accumulation order), so both produce the same bits;
// tests/test_gpu_parity.py::test_fast_tier_matches_general_kernel, ::test_fast_tier_edge_columns.
//
// Two instantiations (lars.cu::launch_fast / launch_fast_gsm):
//   k > 128 (NA = 8 / 16 atoms per lane): 64 slots, the first SPLIT columns of the factor in shared memory and the rest in an
//       L2-resident tail, Gram rows through L1 (zero-padded global copy);
//   64 < k <= 128 (NA = 4, GSM): 32 slots, the whole factor and the Gram in shared memory.
//
// What it is bound by (profiles/r2_lars_fast.md): the L1 data pipe -- LDG, LDS, STS and SHFL all cost wavefronts there, and
// 45 % of them are the Gram rows of the correlation pass (4 bytes per multiply-add, the algorithmic minimum).  Hence:
//   arg-max of the inactive covariances   FMNMX3 + 2 REDUX (no memory traffic)
//   join: g = G64[j, A] (one gather), t = V^T g, u = V t over the packed FP64 inverse factor read in 16-byte pairs; |t|^2 is
//         accumulated inside the second sweep (every lane reads all of t), 1^T u is the one shuffle reduction of a knot
//   equiangular weights (incremental), normalisation, slot table (two entries per LDS.128)
//   correlation pass G[:, A] w: rows of occupied slots only, 16-byte loads UQ rows deep, packed FFMA2
//   step length: packed FADD2 / FMUL2, MUFU.RCP, one unsigned minimum per pair, REDUX min
//   a drop (one per ~30 knots) runs the general kernel's Givens downdate

// a warp-uniform decision, stated so that the compiler can see it (a vote result is uniform by construction): keeps the
// *_sync collectives below it out of the divergent-code trampolines
#define UNI(cond) __any_sync(0xffffffffu, (cond))

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// loads / stores that leave L1 to the Gram rows: the FP64 Gram entries of a join (a 2 KB row touched once) and the
// streamed covariance / code rows do not allocate there
template <bool NA_>
__device__ __forceinline__ double ld_f64_stream(const double* p) {
  if (!NA_) return *p;
  double v;
  asm("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <bool NA_>
__device__ __forceinline__ float2 ld_f32x2_stream(const float2* p) {
  if (!NA_) return *p;
  float2 v;
  asm("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_global_vec(const float2* p) {
  float2 v;
  asm("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
// one lane's piece of a Gram row per load: LW consecutive atoms (a warp-wide load covers 32 LW atoms = LW / 32 KB)
template <int LW> struct RowVec;
template <> struct RowVec<4> { typedef float4 type; };
template <> struct RowVec<2> { typedef float2 type; };
__device__ __forceinline__ void row_zero(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void row_zero(float2& v) { v = make_float2(0.f, 0.f); }
// corr[pairs of this vector] += g * w
__device__ __forceinline__ void row_fma(float2* c2, const float4& g, const float2& ww) {
  c2[0] = ffma2(make_float2(g.x, g.y), ww, c2[0]);
  c2[1] = ffma2(make_float2(g.z, g.w), ww, c2[1]);
}
__device__ __forceinline__ void row_fma(float2* c2, const float2& g, const float2& ww) { c2[0] = ffma2(g, ww, c2[0]); }

// Packed storage of the upper-triangular inverse factor V with columns padded to an even number of entries, so that
// both sweeps read 16-byte pairs: column i holds V[0..i][i] (+ one zero pad entry when i is even) at offset fast_cpad(i).
__host__ __device__ constexpr int fast_cpad(int i) { return ((i + 1) >> 1) * ((i >> 1) + 1) * 2; }
// shared-memory words of one warp: V columns 0..SPLIT-1, g and t (FP64), the slot table (atom, weight)
template <int SMAX, int SPLIT>
__host__ __device__ constexpr int fast_group_words() { return 2 * fast_cpad(SPLIT) + 4 * SMAX + 2 * SMAX; }
template <int SMAX, int SPLIT>
__host__ __device__ constexpr int fast_tail_doubles() { return fast_cpad(SMAX) - fast_cpad(SPLIT); }

template <int NA, int SMAX, int SPLIT, int NT, int UQ, int LW = 4, bool NOAL = false, bool GSM = false>
__global__ void __launch_bounds__(NT, 1) lars_fast_kernel(LarsParams<float> P) {
  typedef float T;
  constexpr int SA = SMAX / 32;          // slot registers per lane (slot p = l + 32 m)
  typedef typename RowVec<LW>::type RowV;
  constexpr int NV = NA / LW;            // loads per Gram row per lane
  constexpr int PV = LW / 2;             // atom pairs per load
  constexpr int KP = 32 * NA;            // padded row length of the Gram copy
  static_assert(SMAX % 32 == 0 && NA % LW == 0 && NA % 2 == 0 && SPLIT % 2 == 0 && SPLIT <= SMAX, "bad tile shape");

  if (P.hint != nullptr && *P.hint != (unsigned)P.run_if) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int k = P.k;
  int l = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  asm volatile("" : "+r"(l));            // (opaque: keeps the lane index in a register instead of re-reading %tid in the loops)
  // GSM (k <= 128): the padded fp32 Gram is staged once per CTA in shared memory (as the general kernel does), rows are LDS.128
  T* Gs = reinterpret_cast<T*>(smem_raw);
  const size_t g_bytes = GSM ? round_up<size_t>((size_t)k * KP * sizeof(T), 128) : 0;
  uint32_t* gbase = reinterpret_cast<uint32_t*>(smem_raw + g_bytes) + (size_t)warp * fast_group_words<SMAX, SPLIT>();
  constexpr int MSMP = fast_cpad(SPLIT);                                           // doubles of V kept in shared memory
  double* Mg = reinterpret_cast<double*>(gbase);                                   // columns 0..SPLIT-1 of V
  double* Mx = P.Mhyb + ((size_t)blockIdx.x * (blockDim.x / 32) + warp) * (size_t)fast_tail_doubles<SMAX, SPLIT>();   // columns >= SPLIT
  double* gs = Mg + MSMP;                                                          // g = G[A, j]
  double* us = gs + SMAX;                                                          // t = V^T g
  SlotW<T>* sw_ = reinterpret_cast<SlotW<T>*>(us + SMAX);                          // (atom, weight) by slot, read in pairs by the correlation pass; free slots (0, 0)
  // stale entries are read (times an exact zero) by the paired sweeps: they must be finite
  for (int i = l; i < MSMP + 2 * SMAX; i += 32) Mg[i] = 0.0;

  auto atom_of = [&](int m) -> int { return ((m / LW) * 32 + l) * LW + (m % LW); };
  unsigned long long Grl = reinterpret_cast<unsigned long long>(P.Gp) + (unsigned long long)l * sizeof(RowV);
  asm volatile("" : "+l"(Grl));
  if (GSM) {
    for (int idx = threadIdx.x; idx < k * KP; idx += blockDim.x) {
      const int a = idx / KP, i = idx - a * KP;
      Gs[idx] = (i < k) ? (T)P.G64[(size_t)a * k + i] : T(0);
    }
    __syncthreads();
  }
  // piece v of this lane's part of Gram row a
  auto row_piece = [&](int a, int v) -> RowV {
    if (GSM) return *reinterpret_cast<const RowV*>(Gs + (size_t)a * KP + (v * 32 + l) * LW);
    return ld_global_vec(reinterpret_cast<const RowV*>(Grl + (unsigned long long)(unsigned)a * (unsigned)(KP * sizeof(T))) + v * 32);
  };
  const double* __restrict__ G64 = P.G64;

  const T tiny = T(1.1754943508222875e-38);
  const T dT = T(P.d);
  const T eps32 = T(1.1920928955078125e-07) * dT;     // covariance units (see lars_kernel)
  const T amin = P.amin * dT;
  const T NINF = -CUDART_INF_F;
  const T BIG = 3.402823466e+38f;

  int cb[SA];                            // offsets of this lane's own columns
#pragma unroll
  for (int m = 0; m < SA; ++m) cb[m] = fast_cpad(l + 32 * m);
  auto Vld = [&](int idx) -> double { return (idx >= MSMP) ? Mx[idx - MSMP] : Mg[idx]; };
  auto Vst = [&](int idx, double v) { if (idx >= MSMP) Mx[idx - MSMP] = v; else Mg[idx] = v; };

  // t_i = sum_{p <= i} V[p][i] src[p] for this lane's columns i < s.  Two rows per step: the column and the source
  // vector are read as 16-byte pairs (the pad entry of an even column is an exact zero, the source beyond s is finite).
  auto sweep_t = [&](double (&t)[SA], const double* src, int s) {
    const double2* src2 = reinterpret_cast<const double2*>(src);
    const int np = (s + 1) >> 1;
    if (s <= 32 && s <= SPLIT) {
      const double2* col = reinterpret_cast<const double2*>(Mg + cb[0]);
      const int ie = (l < s) ? l : -1;
#pragma unroll 2
      for (int q = 0; q < np; ++q) {
        const double2 sp = src2[q];
        if (2 * q <= ie) {
          const double2 v = col[q];
          t[0] += v.x * sp.x;
          t[0] += v.y * sp.y;
        }
      }
    } else if (s <= SPLIT) {
      int ie[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) { const int i = l + 32 * m; ie[m] = (i < s) ? i : -1; }
#pragma unroll 2
      for (int q = 0; q < np; ++q) {
        const double2 sp = src2[q];
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (2 * q <= ie[m]) {
            const double2 v = reinterpret_cast<const double2*>(Mg + cb[m])[q];
            t[m] += v.x * sp.x;
            t[m] += v.y * sp.y;
          }
      }
    } else {
      int ie[SA];
      const double2* colp[SA];                   // generic pointers: a lane's column is in shared or in global memory
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        const int i = l + 32 * m;
        ie[m] = (i < s) ? i : -1;
        colp[m] = reinterpret_cast<const double2*>((cb[m] >= MSMP) ? (Mx + (cb[m] - MSMP)) : (Mg + cb[m]));
      }
#pragma unroll 2
      for (int q = 0; q < np; ++q) {
        const double2 sp = src2[q];
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (2 * q <= ie[m]) {
            const double2 v = colp[m][q];
            t[m] += v.x * sp.x;
            t[m] += v.y * sp.y;
          }
      }
    }
  };
  // u_p = sum_{p <= i < s} V[p][i] src[i].  Two columns per step: the columns i, i + 1 of an even i have the same padded
  // length i + 2 (row i + 1 of column i is the zero pad; src[s] = 0 when s is odd, the column behind it is finite).
  // *ssq (when given) receives sum_i src[i]^2, ascending (the general kernel's order): |t|^2 of a join without a reduction
  auto sweep_u = [&](double (&u)[SA], const double* src, int s, double* ssq = nullptr) {
    double sq = 0.0;
    const double2* src2 = reinterpret_cast<const double2*>(src);
    const int nS = s > SPLIT ? SPLIT : s;
    const int nS1 = nS < 32 ? nS : 32;
    const double* rp = Mg + l;
    int i = 0;
#pragma unroll 2
    for (; i < nS1; i += 2) {
      const double2 sp = src2[i >> 1];
      sq = fma(sp.x, sp.x, sq);
      sq = fma(sp.y, sp.y, sq);
      if (l <= i + 1) {
        u[0] += rp[0] * sp.x;
        u[0] += rp[i + 2] * sp.y;
      }
      rp += 2 * (i + 2);
    }
#pragma unroll 2
    for (; i < nS; i += 2) {
      const double2 sp = src2[i >> 1];
      sq = fma(sp.x, sp.x, sq);
      sq = fma(sp.y, sp.y, sq);
#pragma unroll
      for (int m = 0; m < SA; ++m)
        if (l + 32 * m <= i + 1) {
          u[m] += rp[32 * m] * sp.x;
          u[m] += rp[i + 2 + 32 * m] * sp.y;
        }
      rp += 2 * (i + 2);
    }
    if (s > SPLIT) {
      rp = Mx + l;
#pragma unroll 2
      for (i = SPLIT; i < s; ++i) {
        const double si = src[i];
        sq = fma(si, si, sq);
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (l + 32 * m <= i) u[m] += rp[32 * m] * si;
        rp += (i + 2) & ~1;
      }
    }
    if (ssq) *ssq = sq;
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

    // active atoms and the padding beyond k carry cov = -inf (see lars_kernel); pairs of atoms share a 64-bit register pair
    float2 cov2[NA / 2];
    if (k == KP) {
#pragma unroll
      for (int v = 0; v < NA / 2; ++v)
        cov2[v] = ld_f32x2_stream<NOAL>(reinterpret_cast<const float2*>(crow + ((v / PV) * 32 + l) * LW + (v % PV) * 2));
    } else {
#pragma unroll
      for (int m = 0; m < NA; m += 2) {
        const int i = atom_of(m);
        cov2[m / 2] = make_float2((i < k) ? crow[i] : NINF, (i + 1 < k) ? crow[i + 1] : NINF);
      }
    }
    auto covr = [&](int m) -> T& { return (m & 1) ? cov2[m >> 1].y : cov2[m >> 1].x; };
    // per-slot state lives in the registers of the slot's lane (slot p = l + 32 m): atom, coefficient, weights
    T coef[SA], prev[SA];
    double wd[SA];
    int a_reg[SA];
#pragma unroll
    for (int m = 0; m < SA; ++m) {
      coef[m] = T(0); prev[m] = T(0); wd[m] = 0.0; a_reg[m] = 0;
      SlotW<T> z; z.atom = 0; z.w = T(0);
      sw_[l + 32 * m] = z;
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
      for (int m = 0; m < NA; ++m) best = fmaxf(best, covr(m));
      best = gmaxval<32>(best, 0xffffffffu);
      int bi = 0x7fffffff;
#pragma unroll
      for (int m = NA - 1; m >= 0; --m)
        if (covr(m) == best) bi = atom_of(m);
      bi = __reduce_min_sync(0xffffffffu, bi);
      if (UNI(banned >= 0)) {
        // the atom dropped by the last drop step cannot be the joiner of the first join knot after it (see lars_kernel)
        if (UNI(!drop && bi == banned)) {
          T b2 = NINF;
          int i2 = 0x7fffffff;
#pragma unroll
          for (int m = 0; m < NA; ++m) {
            const T cv = (atom_of(m) == banned) ? NINF : covr(m);
            b2 = cv > b2 ? cv : b2;
          }
          b2 = gmaxval<32>(b2, 0xffffffffu);
#pragma unroll
          for (int m = NA - 1; m >= 0; --m)
            if (covr(m) == b2 && atom_of(m) != banned) i2 = atom_of(m);
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
      if (UNI(n_iter >= P.max_iter)) { handoff = true; break; }

      // ---- 2. atom j joins slot n_act ----
      if (UNI(!drop)) {
        if (UNI(n_act >= SMAX)) { handoff = true; slots_full = true; break; }
        const int j = bi;
        const int s = n_act;
        const double* __restrict__ g64row = G64 + (size_t)j * k;      // row j of the (bitwise symmetric) FP64 Gram
        const double gjj = ld_f64_stream<NOAL>(g64row + j);
        double t[SA], u[SA];
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          t[m] = 0.0; u[m] = 0.0;
          if (32 * m < s) {
            const int p = l + 32 * m;
            double gv = 0.0;
            if (p < s) gv = ld_f64_stream<NOAL>(g64row + a_reg[m]);
            gs[p] = gv;
          }
        }
        __syncwarp();
        sweep_t(t, gs, s);
#pragma unroll
        for (int m = 0; m < SA; ++m)
          if (32 * m < s) us[l + 32 * m] = t[m];      // (zero beyond s: the paired sweep reads us[s] when s is odd)
        __syncwarp();
        double tt = 0.0;
        sweep_u(u, us, s, &tt);                        // |t|^2 comes with the sweep: every lane reads all of t
        double su = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) su += u[m];
        su = gsum<32>(su);                             // 1^T u (a shuffle is an L1-pipe wavefront: one reduction, not two)
        const double sig = gjj - tt;
        double asig = fabs(sig);
        asig = asig > 4.930380657631324e-32 ? asig : 4.930380657631324e-32;
        if (UNI(asig < 1e-14)) { handoff = true; break; }          // degenerate regressor: the general kernel handles it
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (atom_of(m) == j) covr(m) = NINF;
        const double rs = fast_rsqrt(asig);
        const double tau = (1.0 - su) * rs * rs;
        const int cs = fast_cpad(s);
        const int top = s | 1;                         // an even column carries one zero pad entry
        double* vcol = (s < SPLIT) ? (Mg + cs) : (Mx + (cs - MSMP));
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          if (p <= top) vcol[p] = (p < s) ? -u[m] * rs : (p == s ? rs : 0.0);
          if (p <= s) wd[m] = (p == s) ? tau : wd[m] - tau * u[m];
          if (p == s) { a_reg[m] = j; sw_[p].atom = j; }
        }
        sw += tau * (1.0 - su);
        n_act = s + 1;
        max_act = n_act > max_act ? n_act : max_act;
        __syncwarp();
      }

      // the Gram rows of the first batch belong to slots 0..UQ-1, final once the join is done: request them now.  Only the
      // rows of occupied slots are loaded (the L1 data pipe is what this kernel is bound by); the others read as zeros.
      RowV gv0[UQ][NV];
#pragma unroll
      for (int t = 0; t < UQ; ++t) {
        const int a0 = sw_[t].atom;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (t < n_act) gv0[t][v] = row_piece(a0, v);
          else row_zero(gv0[t][v]);
        }
      }
      // ---- 3. normalise the equiangular weights ----
      const double AAd = fast_rsqrt1(sw);
      const T AA = (T)AAd;
      // "alpha increasing" (sklearn _least_angle.py:752-765) or a numerically singular active block (1^T G_AA^-1 1 not
      // positive and finite: AA is then NaN, inf or 0): the general kernel decides
      if (UNI((n_iter > 0 && a_prev < C) || !(AA > T(0) && AA < BIG))) { handoff = true; break; }
      T w[SA];
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        w[m] = (T)(wd[m] * AAd);
        if (32 * m < n_act) {
          const int p = l + 32 * m;
          if (p < n_act) sw_[p].w = w[m];
        }
      }
      __syncwarp();

      // ---- 4. correlation of every atom with the equiangular direction: rows of the active atoms, slot order ----
      float2 corr2[NA / 2];
#pragma unroll
      for (int m = 0; m < NA / 2; ++m) corr2[m] = make_float2(0.f, 0.f);
      {
        SlotW<T> e[UQ];
#pragma unroll
        for (int t = 0; t < UQ; t += 2) {
          const float4 e2 = *reinterpret_cast<const float4*>(sw_ + t);          // two slot entries per load
          e[t].w = e2.y; e[t + 1].w = e2.w;
        }
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
          const float2 ww = make_float2(e[t].w, e[t].w);
#pragma unroll
          for (int v = 0; v < NV; ++v) row_fma(&corr2[PV * v], gv0[t][v], ww);
        }
      }
      for (int q0 = UQ; q0 < n_act; q0 += UQ) {
        const int nr = n_act - q0;
        SlotW<T> e[UQ];
#pragma unroll
        for (int t = 0; t < UQ; t += 2) {
          const float4 e2 = *reinterpret_cast<const float4*>(sw_ + q0 + t);
          e[t].atom = __float_as_int(e2.x); e[t].w = e2.y; e[t + 1].atom = __float_as_int(e2.z); e[t + 1].w = e2.w;
        }
        RowV gv[UQ][NV];
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (t < nr) gv[t][v] = row_piece(e[t].atom, v);
            else row_zero(gv[t][v]);
          }
        }
#pragma unroll
        for (int t = 0; t < UQ; ++t) {
          const float2 ww = make_float2(e[t].w, e[t].w);
#pragma unroll
          for (int v = 0; v < NV; ++v) row_fma(&corr2[PV * v], gv[t][v], ww);
        }
      }

      // ---- 5. step length: min over the candidates v = max(C - cov, 0) / (AA - corr + tiny) with a positive denominator.
      // v >= +0 exactly when the denominator is positive (the numerator is >= +0), and the bit patterns of non-negative
      // floats order like unsigned integers (negative values, -0 and NaN sort above every positive finite one), so the
      // test "den > 0 and v < g1" of the general kernel is one unsigned minimum ----
      unsigned g1u = __float_as_uint(BIG);
      {
        const float2 mone = make_float2(-1.f, -1.f), AA2 = make_float2(AA, AA), C2 = make_float2(C, C), tiny2 = make_float2(tiny, tiny);
#pragma unroll
        for (int m = 0; m < NA / 2; ++m) {
          const float2 den = __fadd2_rn(ffma2(corr2[m], mone, AA2), tiny2);
          float2 num = ffma2(cov2[m], mone, C2);
          num.x = fmaxf(num.x, 0.f);
          num.y = fmaxf(num.y, 0.f);
          float2 r;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(den.x));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(den.y));
          const float2 v = __fmul2_rn(num, r);
          g1u = min(g1u, min(__float_as_uint(v.x), __float_as_uint(v.y)));
        }
      }
      const T g1 = __uint_as_float(__reduce_min_sync(0xffffffffu, g1u));
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
      // ---- 6. move along the path ----
      ++n_iter;
      a_prev = C;
      ghost_atom = -1;
#pragma unroll
      for (int m = 0; m < SA; ++m) {
        prev[m] = coef[m];
        coef[m] = prev[m] + gamma * w[m];
      }
      {
        const float2 mg = make_float2(-gamma, -gamma);
#pragma unroll
        for (int m = 0; m < NA / 2; ++m) cov2[m] = ffma2(corr2[m], mg, cov2[m]);
      }
      kn_s += (unsigned)n_act;
      kn_s2 += (unsigned)(n_act * n_act);

      // ---- 7. atom leaves slot p0 (Givens downdate of the factor; see lars_kernel) ----
      if (drop) {
        const int p0 = dslot;
        int a_d = 0;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int v = __shfl_sync(0xffffffffu, a_reg[m], dslot & 31);
          if ((dslot >> 5) == m) a_d = v;
        }
        const int sD = n_act;
        double rr[SA], mv[SA], pre[SA];
        T gp = T(0);
        double wp0 = 0.0, rp0 = 0.0, mpp = 0.0;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int i = l + 32 * m;
          rr[m] = (i >= p0 && i < n_act) ? Vld(cb[m] + p0) : 0.0;
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
          const int c0 = fast_cpad(p0);
#pragma unroll
          for (int m = 0; m < SA; ++m) {
            const int p = l + 32 * m;
            R[m] = (p <= p0) ? Vld(c0 + p) : 0.0;
            nrow[m] = (p < p0) ? p : (p == p0 ? -1 : p - 1);
          }
          __syncwarp();
          for (int i = p0; i + 1 < sD; ++i) {
            const double c = gs[i + 1], sn = us[i + 1];
            const int cn = fast_cpad(i + 1), co = fast_cpad(i);
#pragma unroll
            for (int m = 0; m < SA; ++m) {
              if (32 * m <= i + 1) {
                const int p = l + 32 * m;
                if (p <= i + 1) {
                  const double X = Vld(cn + p);
                  const double nc = c * R[m] + sn * X;
                  R[m] = c * X - sn * R[m];
                  if (nrow[m] >= 0) Vst(co + nrow[m], nc);   // (rows 0..i of the new column i; its pad entry stays zero)
                }
              }
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          wd[m] = (p < n_act && p != p0) ? wd[m] - fw * mv[m] : 0.0;
        }
        // close the gap: slot p + 1 -> slot p for p >= p0 (the freed last slot ends up with zeros)
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          const T c_dn = __shfl_down_sync(0xffffffffu, coef[m], 1, 32);
          const T p_dn = __shfl_down_sync(0xffffffffu, prev[m], 1, 32);
          const double w_dn = __shfl_down_sync(0xffffffffu, wd[m], 1, 32);
          const int a_dn = __shfl_down_sync(0xffffffffu, a_reg[m], 1, 32);
          const T c_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? coef[m + 1 < SA ? m + 1 : m] : T(0), 0, 32);
          const T p_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? prev[m + 1 < SA ? m + 1 : m] : T(0), 0, 32);
          const double w_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? wd[m + 1 < SA ? m + 1 : m] : 0.0, 0, 32);
          const int a_wr = __shfl_sync(0xffffffffu, (m + 1 < SA) ? a_reg[m + 1 < SA ? m + 1 : m] : 0, 0, 32);
          if (p >= p0) {
            coef[m] = (l == 31) ? c_wr : c_dn;
            prev[m] = (l == 31) ? p_wr : p_dn;
            wd[m] = (l == 31) ? w_wr : w_dn;
            a_reg[m] = (l == 31) ? a_wr : a_dn;
          }
        }
        --n_act;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          const int p = l + 32 * m;
          if (p >= n_act) a_reg[m] = 0;
          if (p >= p0 && p <= n_act) { SlotW<T> e; e.atom = a_reg[m]; e.w = T(0); sw_[p] = e; }     // slot n_act: free again
        }
        sw = sw - wp0 - fw * (sum_m - mpp);
        ++st_drops;
        __syncwarp();
        // exact covariance of the dropped atom (sklearn _least_angle.py:891)
        T part = T(0);
        const T* grow = (GSM ? Gs : P.Gp) + (size_t)a_d * KP;
#pragma unroll
        for (int m = 0; m < SA; ++m) {
          if (32 * m < n_act) {
            const int p = l + 32 * m;
            if (p < n_act) part += grow[a_reg[m]] * coef[m];
          }
        }
        part = gsum<32>(part);
        const T cnew = crow[a_d] - part;
#pragma unroll
        for (int m = 0; m < NA; ++m)
          if (atom_of(m) == a_d) covr(m) = cnew;
        __syncwarp();
      }
    }  // path loop

    // ---- write the code row / hand the column on ----
    if (UNI(!handoff)) {
      T* hrow = P.Ht + (size_t)col * k;
      if (k == KP) {
#pragma unroll
        for (int v = 0; v < NV; ++v) { RowV z; row_zero(z); *reinterpret_cast<RowV*>(hrow + (v * 32 + l) * LW) = z; }
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
        if (p < n_act) hrow[a_reg[m]] = coef[m];
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
