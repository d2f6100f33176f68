// centralized.cuh -- CentralizedStrategy on the device (strategies/centralized.hpp:18-38): the agents of a
// scenario are stacked into one OCP (MultiAgentProblem::build_global_ocp, multi_agent_problem.hpp:52-127) with
// block-diagonal dynamics, stage / terminal costs summed in block order, concatenated bounds, and -- because
// build_global_ocp drops the agents' analytic callbacks (ocp.hpp:117-135 re-installs the defaults) -- finite
// differences for *every* derivative of the stacked functions.  One CTA solves one scenario start to finish
// (T is 10 in the reference's multi-agent examples; state n_s = A*NX up to 128, controls m_s = A*NU up to 64).
//
// Bit-compatibility with the reference's dense evaluation order is kept while exploiting the structure:
//  * a finite difference of the stacked cost changes one or two agents' terms; the stacked value is re-formed
//    as the same left-to-right sum (prefix up to the first changed agent, then the remaining terms in order),
//    so the rounding noise of the 4-point stencils (finite_differences.hpp:155-171,271-285) is reproduced;
//  * FD Jacobians of block-diagonal dynamics are exactly block diagonal (equal values subtract to 0.0), and a
//    dot product that skips exact-zero terms has the same value, so products with A and B touch one block;
//  * everything else (Q_uu LLT with its retry loop, explicit inverse, gains, value update, aliased
//    symmetrisation, line search, stop test) follows ilqr.hpp:92-271 with k-ascending sums per output element,
//    one thread per element.
//
// The body is written as data-parallel phases `for (idx = tid; idx < n; idx += nthr)` separated by barriers, with
// all cross-phase state in the workspace, so tests/csrc/host_emulation.cpp can run it with tid = 0, nthr = 1.
#pragma once
#include <ctime>

#include "ilqr_core.cuh"

namespace mas_b200 {

// Host builds run a solve with one thread (barriers are nothing), or -- tests/csrc/*_threads_test.cpp, the race check under
// ThreadSanitizer -- with host threads as the threads of a CTA and MAS_HOST_THREADS_SYNC(id, count, all) as their barrier.
#if defined(__CUDA_ARCH__)
#define MAS_CTA_SYNC() __syncthreads()
#elif defined(MAS_HOST_THREADS_SYNC)
#define MAS_CTA_SYNC() MAS_HOST_THREADS_SYNC(0, 0, 0)
#else
#define MAS_CTA_SYNC() ((void)0)
#endif

template <class M>
struct StackedProblem {
  int A, T;
  double dt;
  int has_bounds;
  double lo[M::NU], hi[M::NU];
  double tolerance;
  int max_iterations;
  // per scenario (pointers already offset to the scenario)
  const double* x0;   // [ns]
  const double* prm;  // [A][NPs]   NPs = max(NP,1)
  double *X, *U, *Xt, *Ut;     // [(T+1)*ns], [T*ms] nominal and trial, column t at t*ns
  double *K, *kff;             // [T][ms*ns] (K(i,j) at i + j*ms), [T][ms]
  double* work;                // scratch, layout in StackedWork
  double* fast;                // K, Q_ux (padded columns) and L | inv | K^T Q_uu: shared memory when it fits, else in `work`
  double* out_cost;            // [1 + A]: stacked best_cost, then per-agent costs
  int* out_int;                // iterations, status, reg_retries, alpha_trials
  long long* phase_cycles;     // optional [kNumPhases]: SM cycles per phase of this scenario (diagnostics), or null
  double max_ms;               // time budget of the stacked solve (ilqr.hpp:84-90): integer milliseconds since its start, checked at the
                               // top of every iteration; +inf = none
  int use_dmma;                // opt-in (MAS_B200_CENTRALIZED_DMMA=1): dense gain / value-update products on the fp64 tensor cores
};

// Phases timed by MAS_PHASE (thread 0 of the scenario's CTA, clock64 between barriers).
enum StackedPhase { PH_TERMINAL = 0, PH_FD, PH_QASM, PH_LLT, PH_INVERSE, PH_GAINS, PH_VALUE, PH_ROLLOUT, kNumPhases };
#if defined(__CUDA_ARCH__)
#define MAS_PHASE_BEGIN() long long mas_phase_t0 = (P.phase_cycles && tid == 0) ? clock64() : 0
#define MAS_PHASE(k)                                              \
  do {                                                            \
    if (P.phase_cycles && tid == 0) {                             \
      const long long now = clock64();                            \
      P.phase_cycles[k] += now - mas_phase_t0;                    \
      mas_phase_t0 = now;                                         \
    }                                                             \
  } while (0)
#else
#define MAS_PHASE_BEGIN() ((void)0)
#define MAS_PHASE(k) ((void)0)
#endif

// Offsets into the double workspace of one scenario.
struct StackedWork {
  int ns, ms, A, NX, NU;
  size_t Vx, Vxx, Ab, Bb, lx, lu, lxx, luu, lux, fd_set, fd2, Qx, Qu, Qxx, Qux, Quu, Qreg, AtV, BtV, cb, pref, S5, R5, S6, R6, dx, scal, fast, fast_doubles, total;
  MAS_HD StackedWork(int A_, int NX_, int NU_) : A(A_), NX(NX_), NU(NU_) {
    ns = A * NX;
    ms = A * NU;
    size_t o = 0;
    auto take = [&](size_t n) {
      const size_t r = o;
      o += n;
      return r;
    };
    Vx = take(ns);
    Vxx = take(static_cast<size_t>(ns) * ns);
    Ab = take(static_cast<size_t>(A) * NX * NX);
    Bb = take(static_cast<size_t>(A) * NX * NU);
    lx = take(ns);
    lu = take(ms);
    lxx = take(static_cast<size_t>(ns) * ns);
    luu = take(static_cast<size_t>(ms) * ms);
    lux = take(static_cast<size_t>(ms) * ns);
    fd_set = o - Ab;      // Ab .. lux are contiguous: the derivative arrays of one time step
    fd2 = take(fd_set);   // second copy: step t-1's derivatives are produced while step t is factorised
    Qx = take(ns);
    Qu = take(ms);
    Qxx = take(static_cast<size_t>(ns) * ns);
    Qux = take(static_cast<size_t>(ms) * ns);
    Quu = take(static_cast<size_t>(ms) * ms);
    Qreg = take(static_cast<size_t>(ms) * ms);
    AtV = take(static_cast<size_t>(ns) * ns);
    BtV = take(static_cast<size_t>(ms) * ns);
    // fast scratch (offsets below are relative to StackedProblem::fast): K and Q_ux with columns padded to ms + 1
    // (conflict-free column reads from shared memory), a region that holds L and Q_uu_inv during the
    // factorisation and K^T Q_uu afterwards, then the per-agent cost tables of the FD stencils and the scalars
    // (the inverse's columns are padded to ms + 1 as well: one thread owns a column, and an unpadded stride of
    //  ms doubles would put all 32 columns of a warp in the same bank)
    const size_t lu_inv = static_cast<size_t>(ms) * ms + static_cast<size_t>(ms) * (ms + 1);
    const size_t region = static_cast<size_t>(ns) * ms > lu_inv ? static_cast<size_t>(ns) * ms : lu_inv;
    size_t fo = 2 * static_cast<size_t>(ns) * (ms + 1) + region;
    auto ftake = [&](size_t n) {
      const size_t r = fo;
      fo += n;
      return r;
    };
    cb = ftake(A);
    pref = ftake(A + 1);
    S5 = ftake(static_cast<size_t>(A) * NX * 2);
    R5 = ftake(static_cast<size_t>(A) * NU * 2);
    S6 = ftake(static_cast<size_t>(A) * NX * 2);
    R6 = ftake(static_cast<size_t>(A) * NU * 2);
    dx = ftake(ns);
    scal = ftake(16);
    fast_doubles = fo;
    fast = take(fast_doubles);  // fallback location inside `work` when it does not fit in shared memory
    total = o;
  }
};

// ---- opt-in fp64 tensor-core products (mma.sync.m8n8k4.f64 = DMMA.884) ------------------------------------------------
// The dense products of the stacked gain and value update (ilqr.hpp:185-191 at n_s = 128, m_s = 64: K = -Q_uu^-1 Q_ux,
// K^T Q_uu, and V_xx = Q_xx + K^T Q_ux + Q_ux^T K + (K^T Q_uu) K) are real matrix contractions of inner dimension m_s.
// The default path evaluates every output element as a k-ascending sequence of separately rounded multiplications
// and additions -- the reference's arithmetic, which the parity gate asserts bit for bit.  With use_dmma the same
// products run on the tensor cores: fused multiply-adds, four k at a time, in the hardware's own accumulation order.
// That changes rounding, and on this all-finite-difference configuration a rounding change moves the result as much
// as a one-ulp change of the input does, so the mode is never the default and is excluded from parity;
// tools/centralized_dmma.py reports its speed-up and its deviation next to the reference's own one-ulp band.
// Fragment layout (PTX ISA, m8n8k4 .row.col f64): lane = 4 * g + c;  A[g][c], B[c][g], C/D[g][2c], [g][2c + 1].
#if defined(__CUDACC__)
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// C[rows x cols] (+)= sum over `terms` of A_t * B_t, inner dimension kd (multiple of 4), rows % 8 == 0, cols % 16 == 0.
// Element accessors: A_t(i, k) = a[t][i * a_is[t] + k * a_ks[t]] (times a_sign[t]), B_t(k, j) = b[t][k * b_ks[t] + j * b_js[t]].
// Every warp owns 8 x 16 output blocks (two 8 x 8 tiles that share their A fragments).
struct DmmaTerm {
  const double* a;
  size_t a_is, a_ks;
  double a_sign;
  const double* b;
  size_t b_ks, b_js;
};
template <int NT>
__device__ void dmma_product(const DmmaTerm (&term)[NT], int rows, int cols, int kd, const double* c_in, size_t c_ld, double* out0, size_t ld0, double* out1,
                             size_t ld1, int tid, int nthr) {
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int g = lane >> 2, c = lane & 3;
  const int ti = rows / 8, tj = cols / 16;
  for (int blk = warp; blk < ti * tj; blk += nwarps) {
    const int i0 = (blk % ti) * 8, j0 = (blk / ti) * 16;
    const size_t i = i0 + g, ja = j0 + 2 * c, jb = j0 + 8 + 2 * c;
    double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;
    if (c_in) {
      c00 = c_in[i + ja * c_ld];
      c01 = c_in[i + (ja + 1) * c_ld];
      c10 = c_in[i + jb * c_ld];
      c11 = c_in[i + (jb + 1) * c_ld];
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const double* pa = term[t].a + i * term[t].a_is + c * term[t].a_ks;
      const double* pb0 = term[t].b + c * term[t].b_ks + static_cast<size_t>(j0 + g) * term[t].b_js;
      const double* pb1 = pb0 + 8 * term[t].b_js;
      const double sg = term[t].a_sign;
#pragma unroll 4
      for (int k0 = 0; k0 < kd; k0 += 4) {
        const double av = sg * pa[k0 * term[t].a_ks];
        dmma884(c00, c01, av, pb0[k0 * term[t].b_ks]);
        dmma884(c10, c11, av, pb1[k0 * term[t].b_ks]);
      }
    }
    out0[i + ja * ld0] = c00;
    out0[i + (ja + 1) * ld0] = c01;
    out0[i + jb * ld0] = c10;
    out0[i + (jb + 1) * ld0] = c11;
    if (out1) {
      out1[i + ja * ld1] = c00;
      out1[i + (ja + 1) * ld1] = c01;
      out1[i + jb * ld1] = c10;
      out1[i + (jb + 1) * ld1] = c11;
    }
  }
}
#endif

// Barrier over a warp-aligned group of `count` threads of the CTA (named barrier `id`, 1..15); the whole CTA when
// count == all.  Sequential host emulation: nothing; threaded host test: its barrier hook.
MAS_HD void stacked_group_sync(int id, int count, int all) {
#if defined(__CUDA_ARCH__)
  if (count == all) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
#elif defined(MAS_HOST_THREADS_SYNC)
  if (count == all) MAS_HOST_THREADS_SYNC(0, 0, 0);
  else MAS_HOST_THREADS_SYNC(id, count, all);
#else
  (void)id;
  (void)count;
  (void)all;
#endif
}

enum StackedScalar { SC_COST = 0, SC_MERIT = 1, SC_TRIAL = 2, SC_REG = 3, SC_PIVOT = 4, SC_FLAG = 5, SC_INNER = 6, SC_TIMEOUT = 7 };

// nanoseconds on a clock that all threads of the CTA agree on (device: %globaltimer; host emulation: steady_clock)
MAS_HD unsigned long long stacked_now_ns() {
#if defined(__CUDA_ARCH__)
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#else
  return static_cast<unsigned long long>(clock()) * (1000000000ull / CLOCKS_PER_SEC);
#endif
}

// stacked value with agent a's term replaced by va: ((pref[a] + va) + c[a+1]) + ... + c[A-1]
MAS_HD double stacked_sum1(const double* c, const double* pref, int A, int a, double va) {
  double s = pref[a] + va;
  for (int k = a + 1; k < A; ++k) s += c[k];
  return s;
}
// two agents replaced (a != b)
MAS_HD double stacked_sum2(const double* c, const double* pref, int A, int a, double va, int b, double vb) {
  if (a > b) {
    const int ti = a;
    a = b;
    b = ti;
    const double tv = va;
    va = vb;
    vb = tv;
  }
  double s = pref[a] + va;
  for (int k = a + 1; k < b; ++k) s += c[k];
  s += vb;
  for (int k = b + 1; k < A; ++k) s += c[k];
  return s;
}

// N stencil points that replace the same agent's term: N chains of the additions of stacked_sum1, side by side.
template <int N>
MAS_HD void stacked_sum1xN(const double* c, const double* pref, int A, int a, const double* va, double* v) {
  double s[N];
#pragma unroll
  for (int q = 0; q < N; ++q) s[q] = pref[a] + va[q];
  for (int k = a + 1; k < A; ++k) {
    const double ck = c[k];
#pragma unroll
    for (int q = 0; q < N; ++q) s[q] += ck;
  }
#pragma unroll
  for (int q = 0; q < N; ++q) v[q] = s[q];
}

// The four stencil points of a mixed second difference at once: agent a's term takes va[(q >> 1) & 1], agent b's term
// vb[q & 1] (a != b), v[q] = the stacked value.  Every sum performs exactly the additions of stacked_sum2 in the same
// order; computing them side by side gives the fp64 pipe four independent chains instead of one (two before the
// second replaced term, where the pairs still coincide) and loads each c[k] once.
MAS_HD void stacked_sum2x4(const double* c, const double* pref, int A, int a, const double* va, int b, const double* vb, double* v) {
  const bool a_first = a < b;
  const int lo_i = a_first ? a : b, hi_i = a_first ? b : a;
  const double* first = a_first ? va : vb;
  const double* second = a_first ? vb : va;
  double p0 = pref[lo_i] + first[0], p1 = pref[lo_i] + first[1];
  for (int k = lo_i + 1; k < hi_i; ++k) {
    const double ck = c[k];
    p0 += ck;
    p1 += ck;
  }
  // q = 2*ia + ib with ia the variant of agent a and ib the variant of agent b
  double s[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ia = (q >> 1) & 1, ib = q & 1;
    const int f = a_first ? ia : ib, g = a_first ? ib : ia;
    s[q] = (f ? p1 : p0) + second[g];
  }
  for (int k = hi_i + 1; k < A; ++k) {
    const double ck = c[k];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] += ck;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = s[q];
}

// cost of one agent at (x + sx*eps*e_i + ..., u + ...) helpers
template <class M>
MAS_HD double agent_stage_pert(const double* x, const double* u, int t, const double* prm, int ix, double dxv, int jx, double djv, int iu, double duv,
                               int ju, double dju) {
  double xp[M::NX], up[M::NU];
#pragma unroll
  for (int k = 0; k < M::NX; ++k) xp[k] = x[k];
#pragma unroll
  for (int k = 0; k < M::NU; ++k) up[k] = u[k];
  if (ix >= 0) xp[ix] = xp[ix] + dxv;
  if (jx >= 0) xp[jx] = xp[jx] + djv;
  if (iu >= 0) up[iu] = up[iu] + duv;
  if (ju >= 0) up[ju] = up[ju] + dju;
  return M::stage(xp, up, t, prm);
}

// One backward pass over the stacked problem (ilqr.hpp:92-193).  Returns nothing; retries are counted in out_int[2].
// FAST_SHARED: StackedProblem::fast is known to point into shared memory (the device kernel when the scratch fits);
// told to the compiler so that it emits shared-memory loads instead of generic ones.
#if defined(__CUDA_ARCH__)
#define MAS_FAST_IS_SHARED(flag, ptr) \
  do {                                \
    if (flag) __builtin_assume(__isShared(ptr)); \
  } while (0)
#else
#define MAS_FAST_IS_SHARED(flag, ptr) ((void)0)
#endif

template <class M, bool FAST_SHARED = false>
MAS_HD void stacked_backward(const StackedProblem<M>& P, const StackedWork& W, int tid, int nthr) {
  MAS_FAST_IS_SHARED(FAST_SHARED, P.fast);
  constexpr int NX = M::NX, NU = M::NU, NPs = (M::NP > 0 ? M::NP : 1);
  const int A = P.A, ns = W.ns, ms = W.ms, T = P.T;
  double* w = P.work;
  double *Vx = w + W.Vx, *Vxx = w + W.Vxx, *Ab = w + W.Ab, *Bb = w + W.Bb, *lx = w + W.lx, *lu = w + W.lu, *lxx = w + W.lxx, *luu = w + W.luu,
         *lux = w + W.lux, *Qx = w + W.Qx, *Qu = w + W.Qu, *Qxx = w + W.Qxx, *Qux = w + W.Qux, *Quu = w + W.Quu, *Qreg = w + W.Qreg,
         *AtV = w + W.AtV, *BtV = w + W.BtV, *cb = P.fast + W.cb, *pref = P.fast + W.pref, *S5 = P.fast + W.S5,
         *R5 = P.fast + W.R5, *S6 = P.fast + W.S6, *R6 = P.fast + W.R6, *scal = P.fast + W.scal;
  const double e5 = 1e-5, e6 = 1e-6;
  const int ldk = ms + 1;
  const int nthr_all = nthr;
  double* fK = P.fast;                                   // K of the current step, column i at i*ldk
  double* fQ = fK + static_cast<size_t>(ns) * ldk;       // Q_ux, same layout
  double* fC = fQ + static_cast<size_t>(ns) * ldk;
  double *Lm = fC, *inv = fC + static_cast<size_t>(ms) * ms, *KtQ = fC;  // K^T Q_uu reuses the space of L and inv

  MAS_PHASE_BEGIN();
  // ---- terminal value (ilqr.hpp:92-102): FD gradient (eps 1e-6) and Hessian (eps 1e-5) of the stacked terminal cost
  const double* xT = P.X + static_cast<size_t>(T) * ns;
  for (int a = tid; a < A; a += nthr) {
    const double* xa = xT + a * NX;
    const double* pa = P.prm + a * NPs;
    cb[a] = M::terminal(xa, pa);
    for (int i = 0; i < NX; ++i)
      for (int sgn = 0; sgn < 2; ++sgn) {
        double xp[NX];
        for (int k = 0; k < NX; ++k) xp[k] = xa[k];
        xp[i] = sgn ? xa[i] - e5 : xa[i] + e5;
        S5[(a * NX + i) * 2 + sgn] = M::terminal(xp, pa);
        xp[i] = sgn ? xa[i] - e6 : xa[i] + e6;
        S6[(a * NX + i) * 2 + sgn] = M::terminal(xp, pa);
      }
  }
  MAS_CTA_SYNC();
  if (tid == 0) {
    double s = 0.0;
    pref[0] = s;
    for (int a = 0; a < A; ++a) {
      s += cb[a];
      pref[a + 1] = s;
    }
  }
  MAS_CTA_SYNC();
  for (int idx = tid; idx < ns + ns * ns; idx += nthr) {
    if (idx < ns) {
      const int a = idx / NX, il = idx % NX;
      const double fp = stacked_sum1(cb, pref, A, a, S6[(a * NX + il) * 2 + 0]);
      const double fm = stacked_sum1(cb, pref, A, a, S6[(a * NX + il) * 2 + 1]);
      Vx[idx] = MAS_DIV_CONST(fp - fm, 2 * e6);
    } else {
      const int e = idx - ns, i = e % ns, j = e / ns;
      const int a = i / NX, il = i % NX, b = j / NX, jl = j % NX;
      double h;
      if (i == j) {
        const double fp = finite_or_zero(stacked_sum1(cb, pref, A, a, S5[(a * NX + il) * 2 + 0]));
        const double f0 = finite_or_zero(pref[A]);
        const double fm = finite_or_zero(stacked_sum1(cb, pref, A, a, S5[(a * NX + il) * 2 + 1]));
        h = MAS_DIV_CONST(fp - 2 * f0 + fm, e5 * e5);
      } else if (a == b) {
        const double* xa = xT + a * NX;
        const double* pa = P.prm + a * NPs;
        double v[4], cq[4];
        for (int q = 0; q < 4; ++q) {
          double xp[NX];
          for (int k = 0; k < NX; ++k) xp[k] = xa[k];
          xp[il] = (q & 2) ? xa[il] - e5 : xa[il] + e5;
          xp[jl] = (q & 1) ? xa[jl] - e5 : xa[jl] + e5;
          cq[q] = M::terminal(xp, pa);
        }
        stacked_sum1xN<4>(cb, pref, A, a, cq, v);
        for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
        h = MAS_DIV_CONST(v[0] - v[1] - v[2] + v[3], 4 * e5 * e5);
      } else {
        double v[4];
        stacked_sum2x4(cb, pref, A, a, &S5[(a * NX + il) * 2], b, &S5[(b * NX + jl) * 2], v);
        for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
        h = MAS_DIV_CONST(v[0] - v[1] - v[2] + v[3], 4 * e5 * e5);
      }
      Vxx[i + static_cast<size_t>(j) * ns] = h;
    }
  }
  MAS_CTA_SYNC();
  // v_xx = 0.5 * (v_xx + v_xx^T), aliased: lower triangle from old values, then upper from the new lower
  for (int e = tid; e < ns * ns; e += nthr) {
    const int i = e % ns, j = e / ns;
    if (i > j) Vxx[i + static_cast<size_t>(j) * ns] = 0.5 * (Vxx[i + static_cast<size_t>(j) * ns] + Vxx[j + static_cast<size_t>(i) * ns]);
  }
  MAS_CTA_SYNC();
  for (int e = tid; e < ns * ns; e += nthr) {
    const int i = e % ns, j = e / ns;
    if (i <= j) Vxx[i + static_cast<size_t>(j) * ns] = 0.5 * (Vxx[i + static_cast<size_t>(j) * ns] + Vxx[j + static_cast<size_t>(i) * ns]);
  }
  MAS_CTA_SYNC();

  // Two bodies that do not depend on each other inside a step: the finite-difference derivatives of a time step (they
  // need x_t, u_t only) and the factorisation + inverse of Q_uu (64 of the CTA's threads at most, the rest used to wait
  // at its barriers).  With at least four warps, warps 0-1 factorise step t while the other warps already produce
  // step t-1's derivatives into the second set of arrays; otherwise, and in the host emulation, they run in turn.
  const size_t fd_stride = W.fd2 - W.Ab;
  auto fd_phase = [&](int t, int tid, int nthr, int bar_id) {
    const double* xt = P.X + static_cast<size_t>(t) * ns;
    const double* ut = P.U + static_cast<size_t>(t) * ms;
    const size_t off = (t & 1) ? fd_stride : 0;
    double *Ab = w + W.Ab + off, *Bb = w + W.Bb + off, *lx = w + W.lx + off, *lu = w + W.lu + off, *lxx = w + W.lxx + off, *luu = w + W.luu + off,
           *lux = w + W.lux + off;
    auto fsync = [&]() { stacked_group_sync(bar_id, nthr, nthr_all); };
      // ---- per-agent pieces: FD Jacobian blocks, base stage cost, singly perturbed stage costs
      for (int a = tid; a < A; a += nthr) {
        const double* xa = xt + a * NX;
        const double* ua = ut + a * NU;
        const double* pa = P.prm + a * NPs;
        fd_jac_x<M>(xa, ua, pa, Ab + static_cast<size_t>(a) * NX * NX);
        fd_jac_u<M>(xa, ua, pa, Bb + static_cast<size_t>(a) * NX * NU);
        cb[a] = M::stage(xa, ua, t, pa);
      }
      for (int e = tid; e < A * (NX + NU) * 2; e += nthr) {
        const int a = e / ((NX + NU) * 2), r = e % ((NX + NU) * 2), v = r / 2, sgn = r % 2;
        const double* xa = xt + a * NX;
        const double* ua = ut + a * NU;
        const double* pa = P.prm + a * NPs;
        if (v < NX) {
          S5[(a * NX + v) * 2 + sgn] = agent_stage_pert<M>(xa, ua, t, pa, v, sgn ? -e5 : e5, -1, 0, -1, 0, -1, 0);
          S6[(a * NX + v) * 2 + sgn] = agent_stage_pert<M>(xa, ua, t, pa, v, sgn ? -e6 : e6, -1, 0, -1, 0, -1, 0);
        } else {
          const int iu = v - NX;
          R5[(a * NU + iu) * 2 + sgn] = agent_stage_pert<M>(xa, ua, t, pa, -1, 0, -1, 0, iu, sgn ? -e5 : e5, -1, 0);
          R6[(a * NU + iu) * 2 + sgn] = agent_stage_pert<M>(xa, ua, t, pa, -1, 0, -1, 0, iu, sgn ? -e6 : e6, -1, 0);
        }
      }
      fsync();
      if (tid == 0) {
        double s = 0.0;
        pref[0] = s;
        for (int a = 0; a < A; ++a) {
          s += cb[a];
          pref[a + 1] = s;
        }
      }
      fsync();
      // ---- FD derivatives of the stacked stage cost (finite_differences.hpp:110-210,263-287)
      const int n_lx = ns, n_lu = ms, n_lxx = ns * ns, n_luu = ms * ms, n_lux = ms * ns;
      for (int idx = tid; idx < n_lx + n_lu + n_lxx + n_luu + n_lux; idx += nthr) {
        int e = idx;
        if (e < n_lx) {
          const int a = e / NX, il = e % NX;
          double f2[2];
          stacked_sum1xN<2>(cb, pref, A, a, &S6[(a * NX + il) * 2], f2);
          lx[e] = MAS_DIV_CONST(f2[0] - f2[1], 2 * e6);
          continue;
        }
        e -= n_lx;
        if (e < n_lu) {
          const int a = e / NU, il = e % NU;
          double f2[2];
          stacked_sum1xN<2>(cb, pref, A, a, &R6[(a * NU + il) * 2], f2);
          lu[e] = MAS_DIV_CONST(f2[0] - f2[1], 2 * e6);
          continue;
        }
        e -= n_lu;
        if (e < n_lxx + n_luu) {
          const bool is_x = e < n_lxx;
          if (!is_x) e -= n_lxx;
          const int dim = is_x ? ns : ms, per = is_x ? NX : NU;
          const double* tab = is_x ? S5 : R5;
          const int i = e % dim, j = e / dim;
          const int a = i / per, il = i % per, b = j / per, jl = j % per;
          double h;
          if (i == j) {
            double f2[2];
            stacked_sum1xN<2>(cb, pref, A, a, &tab[(a * per + il) * 2], f2);
            const double fp = finite_or_zero(f2[0]);
            const double f0 = finite_or_zero(pref[A]);
            const double fm = finite_or_zero(f2[1]);
            h = MAS_DIV_CONST(fp - 2 * f0 + fm, e5 * e5);
          } else if (a == b) {
            const double* xa = xt + a * NX;
            const double* ua = ut + a * NU;
            const double* pa = P.prm + a * NPs;
            double v[4], cq[4];
            for (int q = 0; q < 4; ++q) {
              const double di = (q & 2) ? -e5 : e5, dj = (q & 1) ? -e5 : e5;
              cq[q] = is_x ? agent_stage_pert<M>(xa, ua, t, pa, il, di, jl, dj, -1, 0, -1, 0)
                           : agent_stage_pert<M>(xa, ua, t, pa, -1, 0, -1, 0, il, di, jl, dj);
            }
            stacked_sum1xN<4>(cb, pref, A, a, cq, v);
            for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
            h = MAS_DIV_CONST(v[0] - v[1] - v[2] + v[3], 4 * e5 * e5);
          } else {
            double v[4];
            stacked_sum2x4(cb, pref, A, a, &tab[(a * per + il) * 2], b, &tab[(b * per + jl) * 2], v);
            for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
            h = MAS_DIV_CONST(v[0] - v[1] - v[2] + v[3], 4 * e5 * e5);
          }
          (is_x ? lxx : luu)[i + static_cast<size_t>(j) * dim] = h;
          continue;
        }
        e -= n_lxx + n_luu;
        {  // cross term H(i,j): control i, state j; f_pp=(x+,u+) f_pm=(x-,u+) f_mp=(x+,u-) f_mm=(x-,u-)
          const int i = e % ms, j = e / ms;
          const int a = i / NU, il = i % NU, b = j / NX, jl = j % NX;
          double v[4];
          if (a == b) {
            double cq[4];
            for (int q = 0; q < 4; ++q) {
              const int sx = q & 1, su = (q >> 1) & 1;  // q: 0 pp, 1 pm (x-), 2 mp (u-), 3 mm
              cq[q] = agent_stage_pert<M>(xt + a * NX, ut + a * NU, t, P.prm + a * NPs, jl, sx ? -e6 : e6, -1, 0, il, su ? -e6 : e6, -1, 0);
            }
            stacked_sum1xN<4>(cb, pref, A, a, cq, v);
            for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
          } else {  // agent a (control, variant su = bit 1 of q), agent b (state, variant sx = bit 0 of q)
            stacked_sum2x4(cb, pref, A, a, &R6[(a * NU + il) * 2], b, &S6[(b * NX + jl) * 2], v);
            for (int q = 0; q < 4; ++q) v[q] = finite_or_zero(v[q]);
          }
          lux[i + static_cast<size_t>(j) * ms] = MAS_DIV_CONST(v[0] - v[1] - v[2] + v[3], 4 * e6 * e6);
        }
      }
    fsync();
  };
  auto llt_inverse_phase = [&](int tid, int nthr, int bar_id) {
    auto lsync = [&]() { stacked_group_sync(bar_id, nthr, nthr_all); };
      // ---- LLT of Q_uu_reg with the cumulative-shift retry loop (ilqr.hpp:172-182); unblocked, lower
      for (;;) {
        // Row i's running sum of squares sq[i] = L_i0^2 + ... (added as the columns are produced, i.e. in the order the
        // reference's pivot computation adds them) makes the pivot of column k a single subtraction; every thread
        // evaluates it redundantly from the untouched diagonal of Q_uu_reg, so a column costs one barrier.
        double* sq = P.fast + W.dx;  // the rollout's scratch, free during the backward pass
        for (int e = tid; e < ms * ms; e += nthr) Lm[e] = Qreg[e];
        for (int i = tid; i < ms; i += nthr) sq[i] = 0.0;
        lsync();
        bool failed = false;
        for (int k = 0; k < ms; ++k) {
          double x = Qreg[k + static_cast<size_t>(k) * ms];
          if (k > 0) x -= sq[k];
          if (x <= 0.0) {
            failed = true;
            break;
          }
          x = sqrt(x);
          if (tid == 0) Lm[k + static_cast<size_t>(k) * ms] = x;
          for (int i = k + 1 + tid; i < ms; i += nthr) {
            double s = Lm[i + static_cast<size_t>(k) * ms];
            if (k > 0) {
              double acc = Lm[i + 0 * static_cast<size_t>(ms)] * Lm[k + 0 * static_cast<size_t>(ms)];
              for (int j = 1; j < k; ++j) acc = acc + Lm[i + static_cast<size_t>(j) * ms] * Lm[k + static_cast<size_t>(j) * ms];
              s -= acc;
            }
            const double l = pm::div_(s, x);
            Lm[i + static_cast<size_t>(k) * ms] = l;
            sq[i] += l * l;
          }
          lsync();
        }
        if (!failed) break;
        const double reg = scal[SC_REG];
        lsync();
        for (int i = tid; i < ms; i += nthr) Qreg[i + static_cast<size_t>(i) * ms] += reg;
        if (tid == 0) {
          scal[SC_REG] = reg * 10.0;
          P.out_int[2] += 1;
        }
        lsync();
        if (!(reg < 1e300)) break;
      }
      // ---- Q_uu_inv = llt.solve(I), one column per thread, the column (stride ldk, conflict-free) is its own work
      // vector.  Forward substitution on e_c: rows above c stay exactly zero, and their products are left out of the
      // later rows' sums (s - L*0 == s).
      for (int c = tid; c < ms; c += nthr) {
        double* x = inv + static_cast<size_t>(c) * ldk;
        for (int i = 0; i < ms; ++i) x[i] = (i == c) ? 1.0 : 0.0;
        for (int i = c; i < ms; ++i) {
          double s = x[i];
          for (int j = c; j < i; ++j) s -= Lm[i + static_cast<size_t>(j) * ms] * x[j];
          x[i] = pm::div_(s, Lm[i + static_cast<size_t>(i) * ms]);
        }
        for (int i = ms - 1; i >= 0; --i) {
          double s = x[i];
          for (int j = i + 1; j < ms; ++j) s -= Lm[j + static_cast<size_t>(i) * ms] * x[j];
          x[i] = pm::div_(s, Lm[i + static_cast<size_t>(i) * ms]);
        }
      }
      lsync();
    lsync();
  };
  const bool overlap = nthr_all >= 128 && (nthr_all % 32) == 0;
  MAS_PHASE(PH_TERMINAL);
  fd_phase(T - 1, tid, nthr, 0);
  MAS_PHASE(PH_FD);
  for (int t = T - 1; t >= 0; --t) {
    const size_t off = (t & 1) ? fd_stride : 0;
    double *Ab = w + W.Ab + off, *Bb = w + W.Bb + off, *lx = w + W.lx + off, *lu = w + W.lu + off, *lxx = w + W.lxx + off, *luu = w + W.luu + off,
           *lux = w + W.lux + off;
    // ---- Q_x, Q_u, A^T V_xx, B^T V_xx (ilqr.hpp:115-119); A, B block diagonal
    for (int idx = tid; idx < ns + ms + ns * ns + ms * ns; idx += nthr) {
      int e = idx;
      if (e < ns) {
        const int a = e / NX, il = e % NX;
        const double* Aa = Ab + static_cast<size_t>(a) * NX * NX;
        double s = Aa[0 + il * NX] * Vx[a * NX + 0];
        for (int k = 1; k < NX; ++k) s = s + Aa[k + il * NX] * Vx[a * NX + k];
        Qx[e] = lx[e] + s;
        continue;
      }
      e -= ns;
      if (e < ms) {
        const int a = e / NU, il = e % NU;
        const double* Ba = Bb + static_cast<size_t>(a) * NX * NU;
        double s = Ba[0 + il * NX] * Vx[a * NX + 0];
        for (int k = 1; k < NX; ++k) s = s + Ba[k + il * NX] * Vx[a * NX + k];
        Qu[e] = lu[e] + s;
        continue;
      }
      e -= ms;
      if (e < ns * ns) {
        const int i = e % ns, j = e / ns, a = i / NX, il = i % NX;
        const double* Aa = Ab + static_cast<size_t>(a) * NX * NX;
        double s = Aa[0 + il * NX] * Vxx[a * NX + 0 + static_cast<size_t>(j) * ns];
        for (int k = 1; k < NX; ++k) s = s + Aa[k + il * NX] * Vxx[a * NX + k + static_cast<size_t>(j) * ns];
        AtV[i + static_cast<size_t>(j) * ns] = s;
        continue;
      }
      e -= ns * ns;
      {
        const int i = e % ms, j = e / ms, a = i / NU, il = i % NU;
        const double* Ba = Bb + static_cast<size_t>(a) * NX * NU;
        double s = Ba[0 + il * NX] * Vxx[a * NX + 0 + static_cast<size_t>(j) * ns];
        for (int k = 1; k < NX; ++k) s = s + Ba[k + il * NX] * Vxx[a * NX + k + static_cast<size_t>(j) * ns];
        BtV[i + static_cast<size_t>(j) * ms] = s;
      }
    }
    MAS_CTA_SYNC();
    // ---- Q_xx = l_xx + (A^T V) A,  Q_ux = l_ux + (B^T V) A,  Q_uu = l_uu + (B^T V) B
    for (int idx = tid; idx < ns * ns + ms * ns + ms * ms; idx += nthr) {
      int e = idx;
      if (e < ns * ns) {
        const int i = e % ns, j = e / ns, b = j / NX, jl = j % NX;
        const double* Aj = Ab + static_cast<size_t>(b) * NX * NX;
        double s = AtV[i + static_cast<size_t>(b * NX + 0) * ns] * Aj[0 + jl * NX];
        for (int k = 1; k < NX; ++k) s = s + AtV[i + static_cast<size_t>(b * NX + k) * ns] * Aj[k + jl * NX];
        Qxx[i + static_cast<size_t>(j) * ns] = lxx[i + static_cast<size_t>(j) * ns] + s;
        continue;
      }
      e -= ns * ns;
      if (e < ms * ns) {
        const int i = e % ms, j = e / ms, b = j / NX, jl = j % NX;
        const double* Aj = Ab + static_cast<size_t>(b) * NX * NX;
        double s = BtV[i + static_cast<size_t>(b * NX + 0) * ms] * Aj[0 + jl * NX];
        for (int k = 1; k < NX; ++k) s = s + BtV[i + static_cast<size_t>(b * NX + k) * ms] * Aj[k + jl * NX];
        Qux[i + static_cast<size_t>(j) * ms] = lux[i + static_cast<size_t>(j) * ms] + s;
        fQ[i + static_cast<size_t>(j) * ldk] = Qux[i + static_cast<size_t>(j) * ms];
        continue;
      }
      e -= ms * ns;
      {
        const int i = e % ms, j = e / ms, b = j / NU, jl = j % NU;
        const double* Bj = Bb + static_cast<size_t>(b) * NX * NU;
        double s = BtV[i + static_cast<size_t>(b * NX + 0) * ms] * Bj[0 + jl * NX];
        for (int k = 1; k < NX; ++k) s = s + BtV[i + static_cast<size_t>(b * NX + k) * ms] * Bj[k + jl * NX];
        Quu[i + static_cast<size_t>(j) * ms] = luu[i + static_cast<size_t>(j) * ms] + s;
        Qreg[i + static_cast<size_t>(j) * ms] = Quu[i + static_cast<size_t>(j) * ms];
      }
    }
    if (tid == 0) scal[SC_REG] = 1e-6;
    MAS_CTA_SYNC();
    MAS_PHASE(PH_QASM);
    if (overlap && t > 0) {
      if (tid < 64) llt_inverse_phase(tid, 64, 1);
      else fd_phase(t - 1, tid - 64, nthr_all - 64, 2);
      MAS_CTA_SYNC();
      MAS_PHASE(PH_LLT);  // with the overlap this slot holds max(LLT + inverse, FD of the next step)
    } else {
      llt_inverse_phase(tid, nthr, 0);
      MAS_PHASE(PH_LLT);
      if (t > 0) fd_phase(t - 1, tid, nthr, 0);
      MAS_PHASE(PH_FD);
    }
    // ---- gains k = (-inv) Q_u, K = (-inv) Q_ux (ilqr.hpp:185-186)
    double* Kt = P.K + static_cast<size_t>(t) * ms * ns;
    double* kt = P.kff + static_cast<size_t>(t) * ms;
    for (int idx = tid; idx < ms; idx += nthr) {
      {
        const int i = idx;
        double s = (-inv[i + 0 * static_cast<size_t>(ldk)]) * Qu[0];
        for (int k = 1; k < ms; ++k) s = s + (-inv[i + static_cast<size_t>(k) * ldk]) * Qu[k];
        kt[i] = s;
      }
    }
    bool dmma = false;
#if defined(__CUDA_ARCH__)
    dmma = P.use_dmma && ns % 16 == 0 && ms % 16 == 0 && nthr % 32 == 0;
    if (dmma) {  // K = (-inv) Q_ux on the tensor cores, into global K_t and the shared copy
      const DmmaTerm tk[1] = {{inv, 1, static_cast<size_t>(ldk), -1.0, fQ, 1, static_cast<size_t>(ldk)}};
      dmma_product<1>(tk, ms, ns, ms, nullptr, 0, Kt, ms, fK, ldk, tid, nthr);
    }
#endif
    for (int e = tid; !dmma && e < ms * ((ns + 3) / 4); e += nthr) {  // K: row i, four columns per thread
      const int i = e % ms, j0 = (e / ms) * 4;
      int jc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) jc[r] = j0 + r < ns ? j0 + r : ns - 1;
      double acc[4];
      {
        const double a0 = -inv[i];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = a0 * fQ[static_cast<size_t>(jc[r]) * ldk];
      }
      for (int k = 1; k < ms; ++k) {
        const double a0 = -inv[i + static_cast<size_t>(k) * ldk];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = acc[r] + a0 * fQ[k + static_cast<size_t>(jc[r]) * ldk];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j0 + r < ns) {
          Kt[i + static_cast<size_t>(j0 + r) * ms] = acc[r];
          fK[i + static_cast<size_t>(j0 + r) * ldk] = acc[r];
        }
    }
    MAS_CTA_SYNC();
    MAS_PHASE(PH_GAINS);
    // ---- K^T Q_uu (unregularised), then the value update (ilqr.hpp:188-192)
#if defined(__CUDA_ARCH__)
    if (dmma) {  // K^T Q_uu
      const DmmaTerm tq[1] = {{fK, static_cast<size_t>(ldk), 1, 1.0, Quu, 1, static_cast<size_t>(ms)}};
      dmma_product<1>(tq, ns, ms, ms, nullptr, 0, KtQ, ns, nullptr, 0, tid, nthr);
    }
#endif
    for (int e = tid; !dmma && e < ns * ((ms + 3) / 4); e += nthr) {  // four columns of K^T Q_uu per thread, as in V_xx below
      const int i = e % ns, j0 = (e / ns) * 4;
      int jc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) jc[r] = j0 + r < ms ? j0 + r : ms - 1;
      double acc[4];
      {
        const double a0 = fK[static_cast<size_t>(i) * ldk];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = a0 * Quu[static_cast<size_t>(jc[r]) * ms];
      }
      // Q_uu lives in global memory (L2 at best: L1 is 20 KB next to 207 KB of shared memory): fetch four rows of
      // the four columns before using any, so that sixteen loads are in flight instead of four
      int k = 1;
      for (; k + 4 <= ms; k += 4) {
        double q[4][4], a0[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          a0[kk] = fK[k + kk + static_cast<size_t>(i) * ldk];
#pragma unroll
          for (int r = 0; r < 4; ++r) q[kk][r] = Quu[k + kk + static_cast<size_t>(jc[r]) * ms];
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r] = acc[r] + a0[kk] * q[kk][r];
      }
      for (; k < ms; ++k) {
        const double a0 = fK[k + static_cast<size_t>(i) * ldk];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = acc[r] + a0 * Quu[k + static_cast<size_t>(jc[r]) * ms];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j0 + r < ms) KtQ[i + static_cast<size_t>(j0 + r) * ns] = acc[r];
    }
    MAS_CTA_SYNC();
    for (int idx = tid; idx < ns; idx += nthr) {
      {
        const int i = idx;
        double t1 = fK[0 + static_cast<size_t>(i) * ldk] * Qu[0];
        for (int k = 1; k < ms; ++k) t1 = t1 + fK[k + static_cast<size_t>(i) * ldk] * Qu[k];
        double t2 = fQ[0 + static_cast<size_t>(i) * ldk] * kt[0];
        for (int k = 1; k < ms; ++k) t2 = t2 + fQ[k + static_cast<size_t>(i) * ldk] * kt[k];
        double t3 = KtQ[i + 0 * static_cast<size_t>(ns)] * kt[0];
        for (int k = 1; k < ms; ++k) t3 = t3 + KtQ[i + static_cast<size_t>(k) * ns] * kt[k];
        Vx[i] = ((Qx[i] + t1) + t2) + t3;
      }
    }
    // V_xx: each thread owns one row index i and four consecutive columns, so that every operand it loads feeds four
    // (K, Q_ux of column i) or three (K, Q_ux of column j) of its twelve independent k-ascending sums
#if defined(__CUDA_ARCH__)
    if (dmma) {  // V_xx = Q_xx + K^T Q_ux + Q_ux^T K + (K^T Q_uu) K, accumulated in that order
      const DmmaTerm tv[3] = {{fK, static_cast<size_t>(ldk), 1, 1.0, fQ, 1, static_cast<size_t>(ldk)},
                              {fQ, static_cast<size_t>(ldk), 1, 1.0, fK, 1, static_cast<size_t>(ldk)},
                              {KtQ, 1, static_cast<size_t>(ns), 1.0, fK, 1, static_cast<size_t>(ldk)}};
      dmma_product<3>(tv, ns, ns, ms, Qxx, ns, Vxx, ns, nullptr, 0, tid, nthr);
    }
#endif
    for (int e = tid; !dmma && e < ns * ((ns + 3) / 4); e += nthr) {
      const int i = e % ns, j0 = (e / ns) * 4;
      int jc[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) jc[r] = j0 + r < ns ? j0 + r : ns - 1;
      double m1[4], m2[4], m3[4];
      {
        const double a1 = fK[static_cast<size_t>(i) * ldk], a2 = fQ[static_cast<size_t>(i) * ldk], a3 = KtQ[i];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double b1 = fQ[static_cast<size_t>(jc[r]) * ldk], b2 = fK[static_cast<size_t>(jc[r]) * ldk];
          m1[r] = a1 * b1;
          m2[r] = a2 * b2;
          m3[r] = a3 * b2;
        }
      }
      for (int k = 1; k < ms; ++k) {
        const double a1 = fK[k + static_cast<size_t>(i) * ldk], a2 = fQ[k + static_cast<size_t>(i) * ldk], a3 = KtQ[i + static_cast<size_t>(k) * ns];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double b1 = fQ[k + static_cast<size_t>(jc[r]) * ldk], b2 = fK[k + static_cast<size_t>(jc[r]) * ldk];
          m1[r] = m1[r] + a1 * b1;
          m2[r] = m2[r] + a2 * b2;
          m3[r] = m3[r] + a3 * b2;
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (j0 + r < ns) {
          const size_t o = i + static_cast<size_t>(j0 + r) * ns;
          Vxx[o] = ((Qxx[o] + m1[r]) + m2[r]) + m3[r];
        }
    }
    MAS_CTA_SYNC();
    // aliased symmetrisation (see symmetrize_aliased): strict lower triangle first, then the rest with the new lower
    // values.  Four entries per trip, all loads before the first store (the stores would otherwise hold back the
    // loads queued behind them, one L2 latency per entry).
    for (int pass = 0; pass < 2; ++pass) {
      for (int e0 = tid * 4; e0 < ns * ns; e0 += nthr * 4) {
        double a[4], b[4];
        bool on[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = e0 + r, i = e % ns, j = e / ns;
          on[r] = e < ns * ns && (pass == 0 ? i > j : i <= j);
          a[r] = on[r] ? Vxx[i + static_cast<size_t>(j) * ns] : 0.0;
          b[r] = on[r] ? Vxx[j + static_cast<size_t>(i) * ns] : 0.0;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (on[r]) Vxx[e0 + r] = 0.5 * (a[r] + b[r]);
      }
      MAS_CTA_SYNC();
    }
    MAS_PHASE(PH_VALUE);
  }
}

// Rollout of the stacked system.  alpha < 0: plain rollout of the controls in Uout (prologue, ilqr.hpp:75-76);
// otherwise the line-search forward pass (ilqr.hpp:208-217) from the nominal (P.X, P.U) with gains.  The stacked
// stage cost is the block-order sum of the agents' costs, accumulated over t.  Result in scal[SC_TRIAL].
template <class M, bool FAST_SHARED = false>
MAS_HD void stacked_rollout(const StackedProblem<M>& P, const StackedWork& W, double alpha, double* Xout, double* Uout, int tid, int nthr) {
  MAS_FAST_IS_SHARED(FAST_SHARED, P.fast);
  constexpr int NX = M::NX, NU = M::NU, NPs = (M::NP > 0 ? M::NP : 1);
  const int A = P.A, ns = W.ns, ms = W.ms, T = P.T;
  double* w = P.work;
  double *cb = P.fast + W.cb, *dxv = P.fast + W.dx, *scal = P.fast + W.scal;
  for (int i = tid; i < ns; i += nthr) Xout[i] = P.x0[i];
  if (tid == 0) scal[SC_TRIAL] = 0.0;
  MAS_CTA_SYNC();
  for (int t = 0; t < T; ++t) {
    double* xt = Xout + static_cast<size_t>(t) * ns;
    double* ut = Uout + static_cast<size_t>(t) * ms;
    if (alpha >= 0.0) {
      const double* xn = P.X + static_cast<size_t>(t) * ns;
      const double* un = P.U + static_cast<size_t>(t) * ms;
      const double* Kt = P.K + static_cast<size_t>(t) * ms * ns;
      const double* kt = P.kff + static_cast<size_t>(t) * ms;
      for (int i = tid; i < ns; i += nthr) dxv[i] = xt[i] - xn[i];
      MAS_CTA_SYNC();
      for (int i = tid; i < ms; i += nthr) {
        double kdx = Kt[i + 0 * static_cast<size_t>(ms)] * dxv[0];
        for (int j = 1; j < ns; ++j) kdx = kdx + Kt[i + static_cast<size_t>(j) * ms] * dxv[j];
        double ui = (un[i] + alpha * kt[i]) + kdx;
        if (P.has_bounds) {
          const double hi = P.hi[i % NU], lo = P.lo[i % NU];
          ui = (hi < ui) ? hi : ui;
          ui = (lo > ui) ? lo : ui;
        }
        ut[i] = ui;
      }
      MAS_CTA_SYNC();
    }
    for (int a = tid; a < A; a += nthr) {
      const double* pa = P.prm + a * NPs;
      cb[a] = M::stage(xt + a * NX, ut + a * NU, t, pa);
      double xn1[NX];
      rk4_step<M>(xt + a * NX, ut + a * NU, pa, P.dt, xn1);
      for (int k = 0; k < NX; ++k) xt[ns + a * NX + k] = xn1[k];
    }
    MAS_CTA_SYNC();
    if (tid == 0) {
      double inner = 0.0;
      for (int a = 0; a < A; ++a) inner += cb[a];
      scal[SC_TRIAL] += inner;
    }
    MAS_CTA_SYNC();
  }
  for (int a = tid; a < A; a += nthr) cb[a] = M::terminal(Xout + static_cast<size_t>(T) * ns + a * NX, P.prm + a * NPs);
  MAS_CTA_SYNC();
  if (tid == 0) {
    double inner = 0.0;
    for (int a = 0; a < A; ++a) inner += cb[a];
    scal[SC_TRIAL] += inner;
  }
  MAS_CTA_SYNC();
}

// The whole centralized solve of one scenario: iLQR::solve on the stacked OCP, then the per-agent cost
// re-evaluation of centralized.hpp:27-36.
template <class M, bool FAST_SHARED = false>
MAS_HD void stacked_solve(const StackedProblem<M>& P, int tid, int nthr) {
  MAS_FAST_IS_SHARED(FAST_SHARED, P.fast);
  constexpr int NX = M::NX, NU = M::NU, NPs = (M::NP > 0 ? M::NP : 1);
  const StackedWork W(P.A, NX, NU);
  const int ns = W.ns, ms = W.ms, T = P.T;
  double* scal = P.fast + W.scal;
  if (tid == 0) {
    P.out_int[0] = 0;
    P.out_int[1] = STATUS_MAX_ITER;
    P.out_int[2] = 0;
    P.out_int[3] = 0;
  }
  MAS_PHASE_BEGIN();
  stacked_rollout<M, FAST_SHARED>(P, W, -1.0, P.X, P.U, tid, nthr);
  MAS_PHASE(PH_ROLLOUT);
  if (tid == 0) {
    scal[SC_COST] = scal[SC_TRIAL];
    scal[SC_MERIT] = scal[SC_TRIAL];
  }
  MAS_CTA_SYNC();
  const bool timed = P.max_ms < 1.7976931348623157e308;
  const unsigned long long start_ns = timed ? stacked_now_ns() : 0ull;
  for (int iter = 0; iter < P.max_iterations; ++iter) {
    if (timed) {  // `elapsed_ms > max_ms -> break`, whole milliseconds, only here (ilqr.hpp:84-90)
      if (tid == 0) scal[SC_TIMEOUT] = static_cast<double>((stacked_now_ns() - start_ns) / 1000000ull) > P.max_ms ? 1.0 : 0.0;
      MAS_CTA_SYNC();
      const bool out_of_time = scal[SC_TIMEOUT] != 0.0;
      MAS_CTA_SYNC();
      if (out_of_time) {
        if (tid == 0) P.out_int[1] = STATUS_TIME_LIMIT;
        break;
      }
    }
    if (tid == 0) P.out_int[0] = iter + 1;
    stacked_backward<M, FAST_SHARED>(P, W, tid, nthr);
#if defined(__CUDA_ARCH__)
    if (P.phase_cycles && tid == 0) mas_phase_t0 = clock64();  // the backward pass accounted for its own phases
#endif
    const double current_merit = scal[SC_MERIT];
    int accepted = -1;
    double best_merit = current_merit;
    double alpha = 1.0;
    for (int j = 0; j < kNumAlphas; ++j) {
      stacked_rollout<M, FAST_SHARED>(P, W, alpha, P.Xt, P.Ut, tid, nthr);
      const double trial = scal[SC_TRIAL];
      MAS_CTA_SYNC();
      MAS_PHASE(PH_ROLLOUT);
      if (tid == 0) P.out_int[3] += 1;
      if (trial < best_merit) {
        best_merit = trial;
        accepted = j;
        break;
      }
      alpha *= 0.5;
    }
    if (accepted >= 0) {
      for (int i = tid; i < (T + 1) * ns; i += nthr) P.X[i] = P.Xt[i];
      for (int i = tid; i < T * ms; i += nthr) P.U[i] = P.Ut[i];
      if (tid == 0) {
        scal[SC_COST] = best_merit;  // objective re-evaluated on the accepted trajectory: the same sum
        scal[SC_MERIT] = best_merit;
      }
    }
    MAS_CTA_SYNC();
    const double improvement = current_merit - best_merit;
    if (improvement < P.tolerance) {
      if (tid == 0) P.out_int[1] = STATUS_CONVERGED;
      break;
    }
  }
  MAS_CTA_SYNC();
  // per-agent best_cost = objective of the agent's own block (centralized.hpp:32), total = stacked best_cost (:27)
  for (int a = tid; a < P.A; a += nthr) {
    const double* pa = P.prm + a * NPs;
    double c = 0.0;
    for (int t = 0; t < T; ++t) c += M::stage(P.X + static_cast<size_t>(t) * ns + a * NX, P.U + static_cast<size_t>(t) * ms + a * NU, t, pa);
    c += M::terminal(P.X + static_cast<size_t>(T) * ns + a * NX, pa);
    P.out_cost[1 + a] = c;
  }
  if (tid == 0) P.out_cost[0] = scal[SC_COST];
}

}  // namespace mas_b200
