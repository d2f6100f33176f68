// ilqr_core.cuh -- per-problem bodies of the batched iLQR kernels (one problem per thread / lane group).
//
// Hot path being replaced: mas::iLQR::solve, include/multi_agent_solver/solvers/ilqr.hpp:59-273, and
// what it calls per iteration: integrate_rk4 / integrate_horizon (integrator.hpp:19-48),
// compute_trajectory_cost (ocp.hpp:14-28), the finite-difference defaults
// (finite_differences.hpp:53-287) and clamp_controls (constraint_helpers.hpp:107-114).
//
// Data layout (HBM, structure of arrays, `ld` = padded batch stride, problem index fastest):
//   x0 [NX][ld]   X [T+1][NX][ld]   U [T][NU][ld]   K [T][NU*NX][ld] (K(i,j) at i + j*NU)   k [T][NU][ld]
// so the 32 lanes of a warp working on 32 neighbouring problems read 32 consecutive doubles.
//
// Arithmetic contract: every sum below is written in the operation order of the reference
// expression it restates (k-ascending dot products starting from the first product, no fused
// multiply-add: compile with -fmad=false).  The functions are __host__ __device__ so the very same
// source can be emulated on the CPU by tests/csrc/host_emulation.cpp.
#pragma once
#include "models.cuh"

namespace mas_b200 {

enum SolveStatus : int { STATUS_CONVERGED = 0, STATUS_MAX_ITER = 1, STATUS_TIME_LIMIT = 2 };

constexpr int kNumAlphas = 10;  // alpha = 1, 1/2, ... while alpha >= 1e-3 (ilqr.hpp:199-206,227)
constexpr int kMaxParams = 8;

template <int NX, int NU>
struct BatchView {
  int ld;
  int T;
  double dt;
  double half_dt, sixth_dt;  // 0.5 * dt and dt / 6 (integrator.hpp:22-27), formed once on the host: IEEE operations, the same bits
  void set_dt(double step) {
    dt = step;
    half_dt = 0.5 * step;
    sixth_dt = step / 6.0;
  }
  unsigned deriv_mask;
  int has_bounds;  // both input bounds present (ilqr.hpp:213)
  double lo[NU], hi[NU];
  double clamp_lo[NU], clamp_hi[NU];  // the same, or -inf / +inf without bounds: the trial rollouts clamp without a branch
  void set_bounds(int has, const double* lower, const double* upper) {
    has_bounds = has;
    for (int i = 0; i < NU; ++i) {
      lo[i] = lower[i];
      hi[i] = upper[i];
      clamp_lo[i] = has ? lower[i] : -__builtin_huge_val();
      clamp_hi[i] = has ? upper[i] : __builtin_huge_val();
    }
  }
  int per_problem_params;  // 0: shared_p, 1: params[NP][ld]
  double shared_p[kMaxParams];
  const double* params;
  const double* x0;
  double* X;
  double* U;
  double* K;
  double* kff;
  double* cost;
  double* merit;
  int* iters;
  int* status;
  int* trials;       // line-search candidates the sequential reference would have evaluated
  int* reg_retries;  // Q_uu + reg*I retries (ilqr.hpp:175-182)
  double tolerance;
  int max_iterations;
  // augmented-Lagrangian state of models with path constraints (ilqr.hpp:31-34,415-418,445-453): one multiplier
  // vector per time step and one penalty parameter per problem; they persist across solves like the members of a
  // reference solver object.  Unused (null) for models without constraints.
  double* lam_eq;    // [T][NEQ][ld]
  double* lam_ineq;  // [T][NINEQ][ld]
  double* penalty;   // [ld]
  double penalty_increase, constraint_tolerance, activation_tolerance;
  // Trial trajectories of the line search, one slot per (thread, chain) of the launch: [T][NX][trial_slots] for
  // x_1..x_T and [T][NU][trial_slots].  The accepted candidate is then copied instead of being rolled out again
  // (a second pass over T dependent RK4 steps).  Null / 0 = not available: the accepted step is recomputed.
  double* trial_X;
  double* trial_U;
  int trial_slots;
  // per-iteration trace of `debug` solves (ilqr.hpp:79-80,262-267): record r of problem p at dbg[(r * kDebugFields + f) * ld + p],
  // r = 0 the initial cost / merit, r = it the values the reference prints after iteration it; null = off
  double* dbg;
  int dbg_records;  // capacity in records (max_iterations + 1)
};

constexpr int kDebugFields = 6;  // cost, merit, d_merit, eq_violation, ineq_violation, accepted step-size index (-1 = none)

constexpr int kMaxALHorizon = 128;  // horizon bound of constrained models (merit addends are kept per step)

// Streaming (evict-first) store / load for data touched once: the trial store must not push the prefetched lines of
// the nominal trajectory out of L1.
MAS_HD void store_streaming(double* p, double v) {
#if defined(__CUDA_ARCH__)
  __stcs(p, v);
#else
  *p = v;
#endif
}
MAS_HD double load_streaming(const double* p) {
#if defined(__CUDA_ARCH__)
  return __ldcs(p);
#else
  return *p;
#endif
}

// Pulls the line holding *p into L1 ahead of use.  Every kernel below walks the trajectory one time
// step at a time with a long dependent fp64 chain per step; asking for step t+1's lines while step
// t computes hides the ~800-cycle HBM latency without spending registers on double buffering.
MAS_HD void prefetch_l1(const void* p) {
#if defined(__CUDA_ARCH__) && !defined(MAS_NO_L1_PREFETCH)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

template <int DIM>
MAS_HD size_t soa_index(int t, int d, int ld, int p) {
  return (static_cast<size_t>(t) * DIM + d) * static_cast<size_t>(ld) + p;
}

// ---- asynchronous staging of the next time step's operands in shared memory -----------------------------------
// `prefetch.global.L1` does not make the demand loads of the next step hit (ncu: 0.003 % L1 hits, 14 % of the line
// search's stall samples on the long scoreboard), so the kernels that walk a trajectory copy step t+1's operands
// into a per-thread shared-memory slot with cp.async while step t computes, and read them back with LDS.
// Layout: stage[(buffer * NV + k) * kStageStride + thread]; a thread only ever touches its own column, so no
// barrier is needed -- cp.async.wait_group orders the thread's own copies.  stage == nullptr (host build, kernels
// without a staging area): plain loads + L1 prefetch.
constexpr int kStageStride = 64;  // threads per CTA of the kernels that stage (engine.cuh: kBlock)
MAS_HD void stage_copy8(double* smem_dst, const double* gmem_src) {
#if defined(__CUDA_ARCH__)
  const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src));
#else
  *smem_dst = *gmem_src;
#endif
}
MAS_HD void stage_commit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;");
#endif
}
MAS_HD void stage_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
// Operands of one step of a trial / commit rollout: x_t (or x_{t+1} for the commit pass), u_t, k_t, K_t.
template <int NX, int NU>
struct StepOperands {
  static constexpr int NV = NX + 2 * NU + NU * NX;
};
template <int NX, int NU>
MAS_HD void stage_issue_step(const BatchView<NX, NU>& v, int p, int t, int t_state, double* stage, int buf) {
  constexpr int NV = StepOperands<NX, NU>::NV;
  double* dst = stage + static_cast<size_t>(buf) * NV * kStageStride;
#pragma unroll
  for (int i = 0; i < NX; ++i) stage_copy8(dst + i * kStageStride, &v.X[soa_index<NX>(t_state, i, v.ld, p)]);
#pragma unroll
  for (int i = 0; i < NU; ++i) stage_copy8(dst + (NX + i) * kStageStride, &v.U[soa_index<NU>(t, i, v.ld, p)]);
#pragma unroll
  for (int i = 0; i < NU; ++i) stage_copy8(dst + (NX + NU + i) * kStageStride, &v.kff[soa_index<NU>(t, i, v.ld, p)]);
#pragma unroll
  for (int i = 0; i < NU * NX; ++i) stage_copy8(dst + (NX + 2 * NU + i) * kStageStride, &v.K[soa_index<NU * NX>(t, i, v.ld, p)]);
  stage_commit();
}
template <int NX, int NU>
MAS_HD void stage_read_step(const double* stage, int buf, double* xn, double* un, double* kv, double* Km) {
  constexpr int NV = StepOperands<NX, NU>::NV;
  const double* src = stage + static_cast<size_t>(buf) * NV * kStageStride;
#pragma unroll
  for (int i = 0; i < NX; ++i) xn[i] = src[i * kStageStride];
#pragma unroll
  for (int i = 0; i < NU; ++i) un[i] = src[(NX + i) * kStageStride];
#pragma unroll
  for (int i = 0; i < NU; ++i) kv[i] = src[(NX + NU + i) * kStageStride];
#pragma unroll
  for (int i = 0; i < NU * NX; ++i) Km[i] = src[(NX + 2 * NU + i) * kStageStride];
}

template <class M, int NX, int NU>
MAS_HD void load_params(const BatchView<NX, NU>& v, int p, double* prm) {
#pragma unroll
  for (int i = 0; i < (M::NP > 0 ? M::NP : 1); ++i) {
    if (i < M::NP) prm[i] = v.per_problem_params ? v.params[static_cast<size_t>(i) * v.ld + p] : v.shared_p[i];
  }
}

// ---- small dense helpers (column-major, compile-time sizes, reference summation order) -----------
// C[R x CC] = A^T * B with A stored KD x R, B stored KD x CC
template <int KD, int R, int CC>
MAS_HD void mat_tn(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < CC; ++j)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      double s = A[0 + i * KD] * B[0 + j * KD];
#pragma unroll
      for (int k = 1; k < KD; ++k) s = s + A[k + i * KD] * B[k + j * KD];
      C[i + j * R] = s;
    }
}
// C[R x CC] = A * B with A stored R x KD, B stored KD x CC
template <int R, int KD, int CC>
MAS_HD void mat_nn(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < CC; ++j)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      double s = A[i + 0 * R] * B[0 + j * KD];
#pragma unroll
      for (int k = 1; k < KD; ++k) s = s + A[i + k * R] * B[k + j * KD];
      C[i + j * R] = s;
    }
}

// The same two products when one factor is a Jacobian with known structural zeros (models.cuh: A_NZ / B_NZ): terms
// whose Jacobian entry is structurally zero are left out, the rest keep the k-ascending order.
// C[R x CC] = A^T * B, A stored KD x R with sparsity NZ (bit k + i*KD)
template <int KD, int R, int CC, unsigned long long NZ>
MAS_HD void mat_tn_sa(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < CC; ++j)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      double s = 0.0;
      bool first = true;
#pragma unroll
      for (int k = 0; k < KD; ++k)
        if ((NZ >> (k + i * KD)) & 1ull) {
          const double prod = A[k + i * KD] * B[k + j * KD];
          s = first ? prod : s + prod;
          first = false;
        }
      C[i + j * R] = s;
    }
}
// C[R x CC] = A * B, B stored KD x CC with sparsity NZ (bit k + j*KD)
template <int R, int KD, int CC, unsigned long long NZ>
MAS_HD void mat_nn_sb(const double* A, const double* B, double* C) {
#pragma unroll
  for (int j = 0; j < CC; ++j)
#pragma unroll
    for (int i = 0; i < R; ++i) {
      double s = 0.0;
      bool first = true;
#pragma unroll
      for (int k = 0; k < KD; ++k)
        if ((NZ >> (k + j * KD)) & 1ull) {
          const double prod = A[i + k * R] * B[k + j * KD];
          s = first ? prod : s + prod;
          first = false;
        }
      C[i + j * R] = s;
    }
}

// `m = 0.5 * (m + m.transpose())` evaluated in place, columns outer / rows inner, as the reference's
// aliased Eigen expression does (ilqr.hpp:102,192).
template <int N>
MAS_HD void symmetrize_aliased(double* m) {
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) m[i + j * N] = 0.5 * (m[i + j * N] + m[j + i * N]);
}

// ---- integrator.hpp:19-28 ---------------------------------------------------------------------
template <class M>
MAS_HD void rk4_step(const double* x, const double* u, const double* prm, double dt, double* xn) {
  constexpr int NX = M::NX;
  double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
  const double hdt = 0.5 * dt;
  double cu[M::NCU];
  M::control_terms(u, prm, cu);  // control-only terms, shared by the four stages
  M::dynamics_c(x, u, cu, prm, k1);
#pragma unroll
  for (int i = 0; i < NX; ++i) xs[i] = x[i] + hdt * k1[i];
  M::dynamics_c(xs, u, cu, prm, k2);
#pragma unroll
  for (int i = 0; i < NX; ++i) xs[i] = x[i] + hdt * k2[i];
  M::dynamics_c(xs, u, cu, prm, k3);
#pragma unroll
  for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt * k3[i];
  M::dynamics_c(xs, u, cu, prm, k4);
  const double sixth = MAS_DIV_CONST(dt, 6.0);
#pragma unroll
  for (int i = 0; i < NX; ++i) xn[i] = x[i] + sixth * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
}

// The same step as one basic block, for the trial rollouts of the line search: every division on its fast path with
// selects instead of branches (portable_math.h: tan_spec, div_const_spec), `hdt` = dt / 2 and `sixth` = dt / 6 from the view
// (constant-bank operands).  Returns false when a division would have needed its slow path (operands outside 2^+-896: trajectories that
// have left the finite range) -- the caller then repeats the step with rk4_step; otherwise xn holds the bits of rk4_step.
// With the branches gone the compiler interleaves the four stages' sin / cos chains and the C candidates of a lane.
template <class M>
MAS_HD bool rk4_step_spec(const double* x, const double* u, const double* prm, double dt, double hdt, double sixth, double* xn) {
#if defined(MAS_NO_SPEC_STEP)  // tuning builds: the branchy step everywhere (A/B measurement of the straight-line step)
  constexpr bool kSpec = false;
#else
  constexpr bool kSpec = M::SPEC_STEP;
#endif
  if constexpr (kSpec) {
    constexpr int NX = M::NX;
    double k1[NX], k2[NX], k3[NX], k4[NX], xs[NX];
    double cu[M::NCU];
    bool exact = true;
    M::control_terms_spec(u, prm, cu, &exact);
    M::dynamics_c_spec(x, u, cu, prm, k1, &exact);
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + hdt * k1[i];
    M::dynamics_c_spec(xs, u, cu, prm, k2, &exact);
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + hdt * k2[i];
    M::dynamics_c_spec(xs, u, cu, prm, k3, &exact);
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[i] + dt * k3[i];
    M::dynamics_c_spec(xs, u, cu, prm, k4, &exact);
#pragma unroll
    for (int i = 0; i < NX; ++i) xn[i] = x[i] + sixth * (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
    return exact;
  } else {
    rk4_step<M>(x, u, prm, dt, xn);
    return true;
  }
}

// ---- prologue: X = integrate_horizon(x0, U); cost = objective(X, U)  (ilqr.hpp:75-78) -----------
// Also used by the strategy layer's re-rollouts (nash.hpp:140,224).  Returns the cost.
template <class M>
MAS_HD double rollout_thread(const BatchView<M::NX, M::NU>& v, int p) {
  constexpr int NX = M::NX, NU = M::NU;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  double x[NX], u[NU], xn[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    x[i] = v.x0[static_cast<size_t>(i) * v.ld + p];
    v.X[soa_index<NX>(0, i, v.ld, p)] = x[i];
  }
  double cost = 0.0;
  for (int t = 0; t < v.T; ++t) {
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    cost += M::stage(x, u, t, prm);
    rk4_step<M>(x, u, prm, v.dt, xn);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      x[i] = xn[i];
      v.X[soa_index<NX>(t + 1, i, v.ld, p)] = x[i];
    }
  }
  cost += M::terminal(x, prm);
  return cost;
}

// ---- finite-difference defaults (finite_differences.hpp) ------------------------------------------
// safe_eval, :95-107 (exponent field all ones = inf or nan; tested with integer instructions, off the fp64 pipe)
MAS_HD double finite_or_zero(double v) { return ((pm::high_word(v) & 0x7ff00000u) != 0x7ff00000u) ? v : 0.0; }

template <class M>
MAS_HD void fd_jac_x(const double* x, const double* u, const double* prm, double* A) {  // :53-72
  constexpr int NX = M::NX;
  const double eps = 1e-6;
  double xp[NX], fp[NX], fm[NX], cu[M::NCU];
  M::control_terms(u, prm, cu);  // u is not perturbed here
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int k = 0; k < NX; ++k) xp[k] = x[k];
    xp[i] = x[i] + eps;
    M::dynamics_c(xp, u, cu, prm, fp);
    xp[i] = x[i] - eps;
    M::dynamics_c(xp, u, cu, prm, fm);
#pragma unroll
    for (int r = 0; r < NX; ++r) A[r + i * NX] = MAS_DIV_CONST(fp[r] - fm[r], 2 * eps);
  }
}
template <class M>
MAS_HD void fd_jac_u(const double* x, const double* u, const double* prm, double* B) {  // :74-92
  constexpr int NX = M::NX, NU = M::NU;
  const double eps = 1e-6;
  double up[NU], fp[NX], fm[NX];
#pragma unroll
  for (int i = 0; i < NU; ++i) {
#pragma unroll
    for (int k = 0; k < NU; ++k) up[k] = u[k];
    up[i] = u[i] + eps;
    M::dynamics(x, up, prm, fp);
    up[i] = u[i] - eps;
    M::dynamics(x, up, prm, fm);
#pragma unroll
    for (int r = 0; r < NX; ++r) B[r + i * NX] = MAS_DIV_CONST(fp[r] - fm[r], 2 * eps);
  }
}
template <class M>
MAS_HD void fd_l_x(const double* x, const double* u, int t, const double* prm, double* g) {  // :110-122
  constexpr int NX = M::NX;
  const double eps = 1e-6;
  double xp[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int k = 0; k < NX; ++k) xp[k] = x[k];
    xp[i] = x[i] + eps;
    const double fp = M::stage(xp, u, t, prm);
    xp[i] = x[i] - eps;
    const double fm = M::stage(xp, u, t, prm);
    g[i] = MAS_DIV_CONST(fp - fm, 2 * eps);
  }
}
template <class M>
MAS_HD void fd_l_u(const double* x, const double* u, int t, const double* prm, double* g) {  // :124-136
  constexpr int NU = M::NU;
  const double eps = 1e-6;
  double up[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) {
#pragma unroll
    for (int k = 0; k < NU; ++k) up[k] = u[k];
    up[i] = u[i] + eps;
    const double fp = M::stage(x, up, t, prm);
    up[i] = u[i] - eps;
    const double fm = M::stage(x, up, t, prm);
    g[i] = MAS_DIV_CONST(fp - fm, 2 * eps);
  }
}
// Hessian of F(z) in z (either the state or the control slot), :138-210 and :229-261.
// F is a functor double(const double* z).
template <int N, class F>
MAS_HD void fd_hessian(const double* z, const F& f, double* H) {
  const double eps = 1e-5;
  double zp[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < N; ++k) zp[k] = z[k];
    zp[i] = z[i] + eps;
    const double fp = finite_or_zero(f(zp));
    const double f0 = finite_or_zero(f(z));
    zp[i] = z[i] - eps;
    const double fm = finite_or_zero(f(zp));
    H[i + i * N] = MAS_DIV_CONST(fp - 2 * f0 + fm, eps * eps);
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
      if (i != j) {
#pragma unroll
        for (int k = 0; k < N; ++k) zp[k] = z[k];
        zp[i] = z[i] + eps;
        zp[j] = z[j] + eps;
        const double fpp = finite_or_zero(f(zp));
        zp[j] = z[j] - eps;
        const double fpm = finite_or_zero(f(zp));
        zp[i] = z[i] - eps;
        zp[j] = z[j] + eps;
        const double fmp = finite_or_zero(f(zp));
        zp[j] = z[j] - eps;
        const double fmm = finite_or_zero(f(zp));
        H[i + j * N] = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * eps * eps);
      }
}
template <class M>
MAS_HD void fd_l_ux(const double* x, const double* u, int t, const double* prm, double* H) {  // :263-287
  constexpr int NX = M::NX, NU = M::NU;
  const double eps = 1e-6;
  double xp[NX], up[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i)
#pragma unroll
    for (int j = 0; j < NX; ++j) {
#pragma unroll
      for (int k = 0; k < NX; ++k) xp[k] = x[k];
#pragma unroll
      for (int k = 0; k < NU; ++k) up[k] = u[k];
      xp[j] = x[j] + eps;
      up[i] = u[i] + eps;
      const double fpp = finite_or_zero(M::stage(xp, up, t, prm));
      xp[j] = x[j] - eps;
      const double fpm = finite_or_zero(M::stage(xp, up, t, prm));
      xp[j] = x[j] + eps;
      up[i] = u[i] - eps;
      const double fmp = finite_or_zero(M::stage(xp, up, t, prm));
      xp[j] = x[j] - eps;
      const double fmm = finite_or_zero(M::stage(xp, up, t, prm));
      H[i + j * NU] = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * eps * eps);
    }
}
template <class M>
struct StageInX {
  const double* u;
  const double* prm;
  int t;
  MAS_HD double operator()(const double* x) const { return M::stage(x, u, t, prm); }
};
template <class M>
struct StageInU {
  const double* x;
  const double* prm;
  int t;
  MAS_HD double operator()(const double* u) const { return M::stage(x, u, t, prm); }
};
template <class M>
struct TerminalInX {
  const double* prm;
  MAS_HD double operator()(const double* x) const { return M::terminal(x, prm); }
};
template <class M>
MAS_HD void fd_v_x(const double* x, const double* prm, double* g) {  // :212-225
  constexpr int NX = M::NX;
  const double eps = 1e-6;
  double xp[NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int k = 0; k < NX; ++k) xp[k] = x[k];
    xp[i] = x[i] + eps;
    const double fp = M::terminal(xp, prm);
    xp[i] = x[i] - eps;
    const double fm = M::terminal(xp, prm);
    g[i] = MAS_DIV_CONST(fp - fm, 2 * eps);
  }
}

// ---- Q_uu regularisation + LLT + explicit inverse (ilqr.hpp:172-183) -----------------------------
// Unblocked lower Cholesky reading the lower triangle only; fails at column k iff the pivot
// x = a_kk - sum_j L_kj^2 is <= 0 (NaN passes), as Eigen::LLT does for these sizes.
template <int N>
MAS_HD bool llt_factor(const double* a, double* L) {
#pragma unroll
  for (int i = 0; i < N * N; ++i) L[i] = a[i];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if (ok) {
      double x = L[k + k * N];
      if (k > 0) {
        double sq = 0.0;
#pragma unroll
        for (int j = 0; j < k; ++j) sq += L[k + j * N] * L[k + j * N];
        x -= sq;
      }
      if (x <= 0.0) {
        ok = false;
      } else {
        x = sqrt(x);
        L[k + k * N] = x;
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
          double s = L[i + k * N];
          if (k > 0) {
            double acc = L[i + 0 * N] * L[k + 0 * N];
#pragma unroll
            for (int j = 1; j < k; ++j) acc = acc + L[i + j * N] * L[k + j * N];
            s -= acc;
          }
          L[i + k * N] = pm::div_(s, x);
        }
      }
    }
  }
  return ok;
}
// inv = A^{-1} via L y = e_c, L^T x = y for each column c of the identity.
template <int N>
MAS_HD void llt_inverse(const double* L, double* inv) {
#pragma unroll
  for (int c = 0; c < N; ++c) {
    double x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double s = x[i];
#pragma unroll
      for (int j = 0; j < i; ++j) s -= L[i + j * N] * x[j];
      x[i] = pm::div_(s, L[i + i * N]);
    }
#pragma unroll
    for (int i = N - 1; i >= 0; --i) {
      double s = x[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) s -= L[j + i * N] * x[j];
      x[i] = pm::div_(s, L[i + i * N]);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) inv[i + c * N] = x[i];
  }
}

// ---- augmented-Lagrangian pieces for models with path constraints -------------------------------------
template <class M>
struct HasConstraints {
  static constexpr bool value = (M::NEQ > 0) || (M::NINEQ > 0);
};

// Constraint Jacobians by central differences, eps = 1e-6 (compute_constraints_state_jacobian /
// _control_jacobian, finite_differences.hpp:289-345).  J is NC x N column-major.
template <class M, bool EQ>
MAS_HD void fd_constraint_jacobians(const double* x, const double* u, const double* prm, double* Jx, double* Ju) {
  constexpr int NX = M::NX, NU = M::NU, NC = EQ ? M::NEQ : M::NINEQ, NCs = NC > 0 ? NC : 1;
  const double eps = 1e-6;
  double xp[NX], up[NU], fp[NCs], fm[NCs];
#pragma unroll
  for (int i = 0; i < NX; ++i) {
#pragma unroll
    for (int k = 0; k < NX; ++k) xp[k] = x[k];
    xp[i] = x[i] + eps;
    if (EQ) M::eq(xp, u, prm, fp);
    else M::ineq(xp, u, prm, fp);
    xp[i] = x[i] - eps;
    if (EQ) M::eq(xp, u, prm, fm);
    else M::ineq(xp, u, prm, fm);
#pragma unroll
    for (int r = 0; r < NC; ++r) Jx[r + i * NC] = MAS_DIV_CONST(fp[r] - fm[r], 2 * eps);
  }
#pragma unroll
  for (int i = 0; i < NU; ++i) {
#pragma unroll
    for (int k = 0; k < NU; ++k) up[k] = u[k];
    up[i] = u[i] + eps;
    if (EQ) M::eq(x, up, prm, fp);
    else M::ineq(x, up, prm, fp);
    up[i] = u[i] - eps;
    if (EQ) M::eq(x, up, prm, fm);
    else M::ineq(x, up, prm, fm);
#pragma unroll
    for (int r = 0; r < NC; ++r) Ju[r + i * NC] = MAS_DIV_CONST(fp[r] - fm[r], 2 * eps);
  }
}

// q += J^T dual;  Q += (rho J_a^T [D]) J_b  with the reference's evaluation order (ilqr.hpp:134-140,158-168):
// first (rho * J_a^T), then (. * D) when `active` is given (a diagonal 0/1 matrix: the off-diagonal products are
// exact zeros), then the product with J_b, every coefficient a k-ascending sum starting from the first product.
template <int NC, int NA, int NB>
MAS_HD void al_add_quadratic(double rho, const double* Ja, const double* Jb, const double* active, double* Q /* NA x NB */) {
#pragma unroll
  for (int j = 0; j < NB; ++j)
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double left = rho * Ja[r + i * NC];
        if (active) left = left * active[r];
        const double term = left * Jb[r + j * NC];
        s = (r == 0) ? term : s + term;
      }
      Q[i + j * NA] = Q[i + j * NA] + s;
    }
}
template <int NC, int N>
MAS_HD void al_add_linear(const double* J, const double* dual, double* q) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = J[0 + i * NC] * dual[0];
#pragma unroll
    for (int r = 1; r < NC; ++r) s = s + J[r + i * NC] * dual[r];
    q[i] = q[i] + s;
  }
}

// Constraint terms of the backward pass at step t (ilqr.hpp:121-170).
template <class M>
MAS_HD void al_backward_terms(const BatchView<M::NX, M::NU>& v, int p, int t, const double* x, const double* u, const double* prm, double rho,
                              double* q_x, double* q_u, double* q_xx, double* q_ux, double* q_uu) {
  constexpr int NX = M::NX, NU = M::NU;
  if (M::NEQ > 0) {
    constexpr int NC = M::NEQ > 0 ? M::NEQ : 1;
    double c[NC], Jx[NC * NX], Ju[NC * NU], dual[NC];
    M::eq(x, u, prm, c);
    fd_constraint_jacobians<M, true>(x, u, prm, Jx, Ju);  // the defaults of ocp.hpp:137-171 ...
    if (v.deriv_mask & D_EQ_JX) M::eq_jac_x(x, u, prm, Jx);  // ... unless the problem installed its own (ocp.hpp:65-68)
    if (v.deriv_mask & D_EQ_JU) M::eq_jac_u(x, u, prm, Ju);
#pragma unroll
    for (int r = 0; r < NC; ++r) dual[r] = v.lam_eq[soa_index<NC>(t, r, v.ld, p)] + rho * c[r];
    al_add_linear<NC, NX>(Jx, dual, q_x);
    al_add_linear<NC, NU>(Ju, dual, q_u);
    al_add_quadratic<NC, NX, NX>(rho, Jx, Jx, nullptr, q_xx);
    al_add_quadratic<NC, NU, NX>(rho, Ju, Jx, nullptr, q_ux);
    al_add_quadratic<NC, NU, NU>(rho, Ju, Ju, nullptr, q_uu);
  }
  if (M::NINEQ > 0) {
    constexpr int NC = M::NINEQ > 0 ? M::NINEQ : 1;
    double g[NC], Jx[NC * NX], Ju[NC * NU], dual[NC], active[NC];
    M::ineq(x, u, prm, g);
    fd_constraint_jacobians<M, false>(x, u, prm, Jx, Ju);
    if (v.deriv_mask & D_INEQ_JX) M::ineq_jac_x(x, u, prm, Jx);
    if (v.deriv_mask & D_INEQ_JU) M::ineq_jac_u(x, u, prm, Ju);
    bool any_active = false;
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      const double slack = g[r] > 0.0 ? g[r] : 0.0;
      active[r] = (g[r] > -v.activation_tolerance) ? 1.0 : 0.0;
      any_active = any_active || active[r] != 0.0;
      dual[r] = v.lam_ineq[soa_index<NC>(t, r, v.ld, p)] * active[r] + rho * slack * active[r];
    }
    al_add_linear<NC, NX>(Jx, dual, q_x);
    al_add_linear<NC, NU>(Ju, dual, q_u);
    if (any_active) {
      al_add_quadratic<NC, NX, NX>(rho, Jx, Jx, active, q_xx);
      al_add_quadratic<NC, NU, NX>(rho, Ju, Jx, active, q_ux);
      al_add_quadratic<NC, NU, NU>(rho, Ju, Ju, active, q_uu);
    }
  }
}

// The up to three addends step t contributes to compute_merit after the objective (ilqr.hpp:386-403).
template <class M>
MAS_HD void al_merit_addends(const BatchView<M::NX, M::NU>& v, int p, int t, const double* x, const double* u, const double* prm, double rho,
                             double* a) {
  a[0] = a[1] = a[2] = 0.0;
  if (M::NEQ > 0) {
    constexpr int NC = M::NEQ > 0 ? M::NEQ : 1;
    double r[NC];
    M::eq(x, u, prm, r);
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      d1 += v.lam_eq[soa_index<NC>(t, i, v.ld, p)] * r[i];
      d2 += r[i] * r[i];
    }
    a[0] = d1 + 0.5 * rho * d2;
  }
  if (M::NINEQ > 0) {
    constexpr int NC = M::NINEQ > 0 ? M::NINEQ : 1;
    double r[NC];
    M::ineq(x, u, prm, r);
    double d1 = 0.0, d2 = 0.0;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const double slack = r[i] > 0.0 ? r[i] : 0.0;
      const double active = (r[i] > -v.activation_tolerance) ? 1.0 : 0.0;
      const double as = slack * active;
      const double w = v.lam_ineq[soa_index<NC>(t, i, v.ld, p)] * active;
      d1 += w * as;
      d2 += as * as;
    }
    a[1] = d1;
    a[2] = 0.5 * rho * d2;
  }
}
// merit = objective, then the per-step addends in time order
template <class M>
MAS_HD double al_finish_merit(double objective, const double* addends, int T) {
  double merit = objective;
  for (int t = 0; t < T; ++t) {
    if (M::NEQ > 0) merit += addends[3 * t + 0];
    if (M::NINEQ > 0) {
      merit += addends[3 * t + 1];
      merit += addends[3 * t + 2];
    }
  }
  return merit;
}

// compute_merit of the trajectory stored in (X, U) (ilqr.hpp:78,380-407)
template <class M>
MAS_HD double al_merit_of_stored(const BatchView<M::NX, M::NU>& v, int p, const double* prm, double objective) {
  constexpr int NX = M::NX, NU = M::NU;
  const double rho = v.penalty[p];
  double merit = objective;
  for (int t = 0; t < v.T; ++t) {
    double x[NX], u[NU], a[3];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    al_merit_addends<M>(v, p, t, x, u, prm, rho, a);
    if (M::NEQ > 0) merit += a[0];
    if (M::NINEQ > 0) {
      merit += a[1];
      merit += a[2];
    }
  }
  return merit;
}

// Multiplier and penalty update after an iteration (ilqr.hpp:236-260) on the stored (X, U); returns the two
// violation norms for the stop test.
template <class M>
MAS_HD void al_update(const BatchView<M::NX, M::NU>& v, int p, const double* prm, double* eq_norm, double* ineq_norm) {
  constexpr int NX = M::NX, NU = M::NU;
  const double rho = v.penalty[p];
  double eq_v = 0.0, ineq_v = 0.0;
  for (int t = 0; t < v.T; ++t) {
    double x[NX], u[NU];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    if (M::NEQ > 0) {
      constexpr int NC = M::NEQ > 0 ? M::NEQ : 1;
      double r[NC];
      M::eq(x, u, prm, r);
      double sq = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        v.lam_eq[soa_index<NC>(t, i, v.ld, p)] += rho * r[i];
        sq += r[i] * r[i];
      }
      eq_v += sq;
    }
    if (M::NINEQ > 0) {
      constexpr int NC = M::NINEQ > 0 ? M::NINEQ : 1;
      double r[NC];
      M::ineq(x, u, prm, r);
      double sq = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const double pos = r[i] > 0.0 ? r[i] : 0.0;
        const double m = v.lam_ineq[soa_index<NC>(t, i, v.ld, p)] + rho * pos;
        v.lam_ineq[soa_index<NC>(t, i, v.ld, p)] = m > 0.0 ? m : 0.0;
        sq += pos * pos;
      }
      ineq_v += sq;
    }
  }
  *eq_norm = sqrt(eq_v);
  *ineq_norm = sqrt(ineq_v);
  if (*eq_norm > v.constraint_tolerance || *ineq_norm > v.constraint_tolerance) v.penalty[p] = rho * v.penalty_increase;
}

// ---- finite-difference derivatives as independent tasks ------------------------------------------------------
// For derivative modes that leave many callbacks to finite differences (all-FD: 118 stage-cost and 12 dynamics
// evaluations per time step at n = 4, m = 2) one thread per problem spends almost all of a step on stencil points
// that do not depend on each other.  The lane-parallel backward pass (engine.cuh: backward_lanes_kernel) deals them
// out as tasks -- one column of A or B, one gradient entry, one Hessian entry -- to the lanes of a problem; each task
// performs exactly the operations the whole-matrix routines above perform for that entry (x - eps is x + (-eps) bit
// for bit), writes its result into the problem's derivative block in shared memory, and lane 0 then runs the
// sequential Riccati step on the assembled block.
template <class M>
struct DerivBlock {  // offsets (doubles) inside a problem's block
  static constexpr int NX = M::NX, NU = M::NU;
  static constexpr int oA = 0, oB = oA + NX * NX, olx = oB + NX * NU, olu = olx + NX, olxx = olu + NU, oluu = olxx + NX * NX, olux = oluu + NU * NU,
                       size = olux + NU * NX;
  // task ranges in the order above: NX columns of A, NU columns of B, NX + NU gradient entries, then the Hessian entries
  static constexpr int tA = 0, tB = tA + NX, tlx = tB + NU, tlu = tlx + NX, tlxx = tlu + NU, tluu = tlxx + NX * NX, tlux = tluu + NU * NU,
                       n_tasks = tlux + NU * NX;
  // terminal value: NX gradient entries and NX*NX Hessian entries, into the first NX + NX*NX doubles of the block
  static constexpr int n_terminal_tasks = NX + NX * NX;
};

// zp = z with d added to component i (and e to component j): a select per component, so that i and j may be run-time
template <int N>
MAS_HD void perturbed(const double* z, int i, double d, int j, double e, double* zp) {
#pragma unroll
  for (int k = 0; k < N; ++k) zp[k] = (k == i) ? z[k] + d : ((k == j) ? z[k] + e : z[k]);
}
// entry (i, j) of fd_hessian<N>(z, f, H)
template <int N, class F>
MAS_HD double fd_hessian_entry(const double* z, const F& f, int i, int j) {
  const double eps = 1e-5;
  double zp[N];
  if (i == j) {
    perturbed<N>(z, i, eps, -1, 0.0, zp);
    const double fp = finite_or_zero(f(zp));
    const double f0 = finite_or_zero(f(z));
    perturbed<N>(z, i, -eps, -1, 0.0, zp);
    const double fm = finite_or_zero(f(zp));
    return MAS_DIV_CONST(fp - 2 * f0 + fm, eps * eps);
  }
  perturbed<N>(z, i, eps, j, eps, zp);
  const double fpp = finite_or_zero(f(zp));
  perturbed<N>(z, i, eps, j, -eps, zp);
  const double fpm = finite_or_zero(f(zp));
  perturbed<N>(z, i, -eps, j, eps, zp);
  const double fmp = finite_or_zero(f(zp));
  perturbed<N>(z, i, -eps, j, -eps, zp);
  const double fmm = finite_or_zero(f(zp));
  return MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * eps * eps);
}
// One task of the stage derivatives; groups whose callback is analytic in `mask` are skipped (lane 0 fills them).
template <class M>
MAS_HD void fd_stage_task(unsigned mask, int task, const double* x, const double* u, int t, const double* prm, double* blk) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  const double e6 = 1e-6;
  if (task < D::tB) {  // column i of A, fd_jac_x
    if (mask & D_A) return;
    const int i = task - D::tA;
    double xp[NX], fp[NX], fm[NX], cu[M::NCU];
    M::control_terms(u, prm, cu);
    perturbed<NX>(x, i, e6, -1, 0.0, xp);
    M::dynamics_c(xp, u, cu, prm, fp);
    perturbed<NX>(x, i, -e6, -1, 0.0, xp);
    M::dynamics_c(xp, u, cu, prm, fm);
#pragma unroll
    for (int r = 0; r < NX; ++r) blk[D::oA + r + i * NX] = MAS_DIV_CONST(fp[r] - fm[r], 2 * e6);
  } else if (task < D::tlx) {  // column i of B, fd_jac_u
    if (mask & D_B) return;
    const int i = task - D::tB;
    double up[NU], fp[NX], fm[NX];
    perturbed<NU>(u, i, e6, -1, 0.0, up);
    M::dynamics(x, up, prm, fp);
    perturbed<NU>(u, i, -e6, -1, 0.0, up);
    M::dynamics(x, up, prm, fm);
#pragma unroll
    for (int r = 0; r < NX; ++r) blk[D::oB + r + i * NX] = MAS_DIV_CONST(fp[r] - fm[r], 2 * e6);
  } else if (task < D::tlu) {  // fd_l_x
    if (mask & D_LX) return;
    const int i = task - D::tlx;
    double xp[NX];
    perturbed<NX>(x, i, e6, -1, 0.0, xp);
    const double fp = M::stage(xp, u, t, prm);
    perturbed<NX>(x, i, -e6, -1, 0.0, xp);
    const double fm = M::stage(xp, u, t, prm);
    blk[D::olx + i] = MAS_DIV_CONST(fp - fm, 2 * e6);
  } else if (task < D::tlxx) {  // fd_l_u
    if (mask & D_LU) return;
    const int i = task - D::tlu;
    double up[NU];
    perturbed<NU>(u, i, e6, -1, 0.0, up);
    const double fp = M::stage(x, up, t, prm);
    perturbed<NU>(u, i, -e6, -1, 0.0, up);
    const double fm = M::stage(x, up, t, prm);
    blk[D::olu + i] = MAS_DIV_CONST(fp - fm, 2 * e6);
  } else if (task < D::tluu) {  // l_xx(i, j)
    if (mask & D_LXX) return;
    const int e = task - D::tlxx, i = e % NX, j = e / NX;
    blk[D::olxx + i + j * NX] = fd_hessian_entry<NX>(x, StageInX<M>{u, prm, t}, i, j);
  } else if (task < D::tlux) {  // l_uu(i, j)
    if (mask & D_LUU) return;
    const int e = task - D::tluu, i = e % NU, j = e / NU;
    blk[D::oluu + i + j * NU] = fd_hessian_entry<NU>(u, StageInU<M>{x, prm, t}, i, j);
  } else {  // l_ux(i, j): control i, state j, fd_l_ux
    if (mask & D_LUX) return;
    const int e = task - D::tlux, i = e % NU, j = e / NU;
    double xp[NX], up[NU];
    perturbed<NX>(x, j, e6, -1, 0.0, xp);
    perturbed<NU>(u, i, e6, -1, 0.0, up);
    const double fpp = finite_or_zero(M::stage(xp, up, t, prm));
    perturbed<NX>(x, j, -e6, -1, 0.0, xp);
    const double fpm = finite_or_zero(M::stage(xp, up, t, prm));
    perturbed<NX>(x, j, e6, -1, 0.0, xp);
    perturbed<NU>(u, i, -e6, -1, 0.0, up);
    const double fmp = finite_or_zero(M::stage(xp, up, t, prm));
    perturbed<NX>(x, j, -e6, -1, 0.0, xp);
    const double fmm = finite_or_zero(M::stage(xp, up, t, prm));
    blk[D::olux + i + j * NU] = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * e6 * e6);
  }
}
// One task of the terminal value: v_x(i) for task < NX, else v_xx(i, j); blk = [v_x | v_xx].
template <class M>
MAS_HD void fd_terminal_task(unsigned mask, int task, const double* x, const double* prm, double* blk) {
  constexpr int NX = M::NX;
  const double e6 = 1e-6;
  if (task < NX) {
    if (mask & D_VX) return;
    double xp[NX];
    perturbed<NX>(x, task, e6, -1, 0.0, xp);
    const double fp = M::terminal(xp, prm);
    perturbed<NX>(x, task, -e6, -1, 0.0, xp);
    const double fm = M::terminal(xp, prm);
    blk[task] = MAS_DIV_CONST(fp - fm, 2 * e6);
  } else {
    if (mask & D_VXX) return;
    const int e = task - NX, i = e % NX, j = e / NX;
    blk[NX + i + j * NX] = fd_hessian_entry<NX>(x, TerminalInX<M>{prm}, i, j);
  }
}
// Lane 0 after the tasks: the derivative arrays of the step, analytic where `mask` says so, else from the block.
template <class M>
MAS_HD void gather_stage_derivatives(unsigned mask, const double* blk, const double* x, const double* u, int t, const double* prm, double* A, double* B,
                                     double* l_x, double* l_u, double* l_xx, double* l_uu, double* l_ux) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  if (mask & D_A) M::jac_x(x, u, prm, A);
  else
    for (int i = 0; i < NX * NX; ++i) A[i] = blk[D::oA + i];
  if (mask & D_B) M::jac_u(x, u, prm, B);
  else
    for (int i = 0; i < NX * NU; ++i) B[i] = blk[D::oB + i];
  if (mask & D_LX) M::l_x(x, u, t, prm, l_x);
  else
    for (int i = 0; i < NX; ++i) l_x[i] = blk[D::olx + i];
  if (mask & D_LU) M::l_u(x, u, t, prm, l_u);
  else
    for (int i = 0; i < NU; ++i) l_u[i] = blk[D::olu + i];
  if (mask & D_LXX) M::l_xx(x, u, t, prm, l_xx);
  else
    for (int i = 0; i < NX * NX; ++i) l_xx[i] = blk[D::olxx + i];
  if (mask & D_LUU) M::l_uu(x, u, t, prm, l_uu);
  else
    for (int i = 0; i < NU * NU; ++i) l_uu[i] = blk[D::oluu + i];
  if (mask & D_LUX) M::l_ux(x, u, t, prm, l_ux);
  else
    for (int i = 0; i < NU * NX; ++i) l_ux[i] = blk[D::olux + i];
}

// ---- one time step of the Riccati recursion (ilqr.hpp:115-193) given the derivatives at (x_t, u_t) --------------
// Q assembly, constraint terms, Q_uu regularisation + LLT + inverse, gains (stored), value update in place.
// Returns the number of regularisation retries of this step.
template <class M, int MASK_CT>
MAS_HD int riccati_step(const BatchView<M::NX, M::NU>& v, int p, int t, const double* x, const double* u, const double* prm, double al_rho,
                        const double* A, const double* B, const double* l_x, const double* l_u, const double* l_xx, const double* l_uu,
                        const double* l_ux, double* v_x, double* v_xx) {
  constexpr int NX = M::NX, NU = M::NU;
  int retries = 0;
  // :115-119
  double q_x[NX], q_u[NU], q_xx[NX * NX], q_ux[NU * NX], q_uu[NU * NU];
  constexpr int TMPN = (NX > NU ? NX : NU) * (NX > NU ? NX : NU);
  double AtV[NX * NX], BtV[NU * NX], tmp[TMPN];
  // structural zeros are only known for the analytic Jacobians of a compile-time derivative mode
  constexpr unsigned long long kDense = ~0ull;
  static_assert(NX * NX <= 64 && NX * NU <= 64, "sparsity masks are 64-bit");
  constexpr unsigned long long a_nz = (MASK_CT >= 0 && (MASK_CT & D_A)) ? M::A_NZ : kDense;
  constexpr unsigned long long b_nz = (MASK_CT >= 0 && (MASK_CT & D_B)) ? M::B_NZ : kDense;
  mat_tn_sa<NX, NX, 1, a_nz>(A, v_x, tmp);
#pragma unroll
  for (int i = 0; i < NX; ++i) q_x[i] = l_x[i] + tmp[i];
  mat_tn_sa<NX, NU, 1, b_nz>(B, v_x, tmp);
#pragma unroll
  for (int i = 0; i < NU; ++i) q_u[i] = l_u[i] + tmp[i];
  mat_tn_sa<NX, NX, NX, a_nz>(A, v_xx, AtV);
  mat_tn_sa<NX, NU, NX, b_nz>(B, v_xx, BtV);
  mat_nn_sb<NX, NX, NX, a_nz>(AtV, A, tmp);
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) q_xx[i] = l_xx[i] + tmp[i];
  mat_nn_sb<NU, NX, NX, a_nz>(BtV, A, tmp);
#pragma unroll
  for (int i = 0; i < NU * NX; ++i) q_ux[i] = l_ux[i] + tmp[i];
  mat_nn_sb<NU, NX, NU, b_nz>(BtV, B, tmp);
#pragma unroll
  for (int i = 0; i < NU * NU; ++i) q_uu[i] = l_uu[i] + tmp[i];

  // :121-170 constraint terms (models with path constraints only)
  if (HasConstraints<M>::value) al_backward_terms<M>(v, p, t, x, u, prm, al_rho, q_x, q_u, q_xx, q_ux, q_uu);

  // :172-183
  double q_reg[NU * NU], L[NU * NU], inv[NU * NU];
#pragma unroll
  for (int i = 0; i < NU * NU; ++i) q_reg[i] = q_uu[i];
  double reg = 1e-6;
  while (!llt_factor<NU>(q_reg, L)) {
#pragma unroll
    for (int i = 0; i < NU; ++i) q_reg[i + i * NU] += reg;
    reg *= 10.0;
    ++retries;
    if (!(reg < 1e300)) break;  // reference loops forever on NaN-free garbage; bail out instead
  }
  llt_inverse<NU>(L, inv);

  // :185-186  k = (-Q_uu_inv) q_u,  K = (-Q_uu_inv) Q_ux
  double ninv[NU * NU], kv[NU], Km[NU * NX];
#pragma unroll
  for (int i = 0; i < NU * NU; ++i) ninv[i] = -inv[i];
  mat_nn<NU, NU, 1>(ninv, q_u, kv);
  mat_nn<NU, NU, NX>(ninv, q_ux, Km);

#pragma unroll
  for (int i = 0; i < NU; ++i) v.kff[soa_index<NU>(t, i, v.ld, p)] = kv[i];
#pragma unroll
  for (int i = 0; i < NU * NX; ++i) v.K[soa_index<NU * NX>(t, i, v.ld, p)] = Km[i];

  // :188-192 value update with the unregularised Q_uu
  double KtQuu[NX * NU], t1[NX], t2[NX], t3[NX];
  mat_tn<NU, NX, NU>(Km, q_uu, KtQuu);
  mat_tn<NU, NX, 1>(Km, q_u, t1);
  mat_tn<NU, NX, 1>(q_ux, kv, t2);
  mat_nn<NX, NU, 1>(KtQuu, kv, t3);
#pragma unroll
  for (int i = 0; i < NX; ++i) v_x[i] = ((q_x[i] + t1[i]) + t2[i]) + t3[i];
  double m1[NX * NX], m2[NX * NX], m3[NX * NX];
  mat_tn<NU, NX, NX>(Km, q_ux, m1);
  mat_tn<NU, NX, NX>(q_ux, Km, m2);
  mat_nn<NX, NU, NX>(KtQuu, Km, m3);
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) v_xx[i] = ((q_xx[i] + m1[i]) + m2[i]) + m3[i];
  symmetrize_aliased<NX>(v_xx);
  return retries;
}

// ---- backward pass for one problem (ilqr.hpp:92-193) --------------------------------------------
// MASK_CT >= 0 fixes the derivative mode at compile time (dead branches vanish); -1 reads it from
// the view.  Writes K, k for every t.  Returns the number of regularisation retries.
// stage != nullptr: x_{t-1}, u_{t-1} are copied to shared memory with cp.async while step t computes.
template <class M, int MASK_CT>
MAS_HD int backward_thread(const BatchView<M::NX, M::NU>& v, int p, double* stage = nullptr) {
  constexpr int NX = M::NX, NU = M::NU;
  const unsigned mask = (MASK_CT >= 0) ? static_cast<unsigned>(MASK_CT) : v.deriv_mask;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  const int T = v.T;
  int retries = 0;
  const double al_rho = HasConstraints<M>::value ? v.penalty[p] : 0.0;

  double x[NX], u[NU];
  double v_x[NX], v_xx[NX * NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(T, i, v.ld, p)];
  // :92-102 terminal value
  if (mask & D_VX) M::v_x(x, prm, v_x);
  else fd_v_x<M>(x, prm, v_x);
  if (mask & D_VXX) M::v_xx(x, prm, v_xx);
  else fd_hessian<NX>(x, TerminalInX<M>{prm}, v_xx);
  symmetrize_aliased<NX>(v_xx);

  constexpr int NVB = NX + NU;  // staged operands per step
  auto issue_xu = [&](int t, int buf) {
    double* dst = stage + static_cast<size_t>(buf) * NVB * kStageStride;
#pragma unroll
    for (int i = 0; i < NX; ++i) stage_copy8(dst + i * kStageStride, &v.X[soa_index<NX>(t, i, v.ld, p)]);
#pragma unroll
    for (int i = 0; i < NU; ++i) stage_copy8(dst + (NX + i) * kStageStride, &v.U[soa_index<NU>(t, i, v.ld, p)]);
    stage_commit();
  };
  if (stage) issue_xu(T - 1, (T - 1) & 1);
  for (int t = T - 1; t >= 0; --t) {
    if (stage) {
      stage_wait();
      const double* src = stage + static_cast<size_t>(t & 1) * NVB * kStageStride;
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = src[i * kStageStride];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = src[(NX + i) * kStageStride];
      if (t > 0) issue_xu(t - 1, (t - 1) & 1);
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
      if (t > 0) {
#pragma unroll
        for (int i = 0; i < NX; ++i) prefetch_l1(&v.X[soa_index<NX>(t - 1, i, v.ld, p)]);
#pragma unroll
        for (int i = 0; i < NU; ++i) prefetch_l1(&v.U[soa_index<NU>(t - 1, i, v.ld, p)]);
      }
    }

    // :106-113
    double A[NX * NX], B[NX * NU], l_x[NX], l_u[NU], l_xx[NX * NX], l_uu[NU * NU], l_ux[NU * NX];
    if (mask & D_A) M::jac_x(x, u, prm, A);
    else fd_jac_x<M>(x, u, prm, A);
    if (mask & D_B) M::jac_u(x, u, prm, B);
    else fd_jac_u<M>(x, u, prm, B);
    if (mask & D_LX) M::l_x(x, u, t, prm, l_x);
    else fd_l_x<M>(x, u, t, prm, l_x);
    if (mask & D_LU) M::l_u(x, u, t, prm, l_u);
    else fd_l_u<M>(x, u, t, prm, l_u);
    if (mask & D_LXX) M::l_xx(x, u, t, prm, l_xx);
    else fd_hessian<NX>(x, StageInX<M>{u, prm, t}, l_xx);
    if (mask & D_LUU) M::l_uu(x, u, t, prm, l_uu);
    else fd_hessian<NU>(u, StageInU<M>{x, prm, t}, l_uu);
    if (mask & D_LUX) M::l_ux(x, u, t, prm, l_ux);
    else fd_l_ux<M>(x, u, t, prm, l_ux);

    retries += riccati_step<M, MASK_CT>(v, p, t, x, u, prm, al_rho, A, B, l_x, l_u, l_xx, l_uu, l_ux, v_x, v_xx);
  }
  return retries;
}

// ---- lane-parallel backward pass: LB lanes per problem --------------------------------------------------------
// Every lane of the group walks the time loop; per step the lanes take the FD tasks lane, lane + LB, ... and write the
// problem's derivative block `blk` (shared memory); after a group barrier lane 0 assembles the derivatives and runs
// riccati_step (V_x, V_xx live in its registers).  `group_sync` is __syncwarp over the group's lanes on the device and
// a no-op in the sequential host emulation, which calls this function once per lane *phase* instead (see
// tests/csrc/host_emulation.cpp).  Returns the regularisation retries (lane 0).
template <class M, int MASK_CT, int LB, class Sync>
MAS_HD int backward_lanes(const BatchView<M::NX, M::NU>& v, int p, int lane, double* blk, const Sync& group_sync) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  const unsigned mask = (MASK_CT >= 0) ? static_cast<unsigned>(MASK_CT) : v.deriv_mask;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  const int T = v.T;
  int retries = 0;
  const double al_rho = HasConstraints<M>::value ? v.penalty[p] : 0.0;
  double x[NX], u[NU], v_x[NX], v_xx[NX * NX];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(T, i, v.ld, p)];
  for (int task = lane; task < D::n_terminal_tasks; task += LB) fd_terminal_task<M>(mask, task, x, prm, blk);
  group_sync();
  if (lane == 0) {
    if (mask & D_VX) M::v_x(x, prm, v_x);
    else
      for (int i = 0; i < NX; ++i) v_x[i] = blk[i];
    if (mask & D_VXX) M::v_xx(x, prm, v_xx);
    else
      for (int i = 0; i < NX * NX; ++i) v_xx[i] = blk[NX + i];
    symmetrize_aliased<NX>(v_xx);
  }
  group_sync();
  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    if (t > 0) {
#pragma unroll
      for (int i = 0; i < NX; ++i) prefetch_l1(&v.X[soa_index<NX>(t - 1, i, v.ld, p)]);
#pragma unroll
      for (int i = 0; i < NU; ++i) prefetch_l1(&v.U[soa_index<NU>(t - 1, i, v.ld, p)]);
    }
    for (int task = lane; task < D::n_tasks; task += LB) fd_stage_task<M>(mask, task, x, u, t, prm, blk);
    group_sync();
    if (lane == 0) {
      double A[NX * NX], B[NX * NU], l_x[NX], l_u[NU], l_xx[NX * NX], l_uu[NU * NU], l_ux[NU * NX];
      gather_stage_derivatives<M>(mask, blk, x, u, t, prm, A, B, l_x, l_u, l_xx, l_uu, l_ux);
      retries += riccati_step<M, MASK_CT>(v, p, t, x, u, prm, al_rho, A, B, l_x, l_u, l_xx, l_uu, l_ux, v_x, v_xx);
    }
    group_sync();
  }
  return retries;
}

// ---- time-parallel linearisation + Riccati sweep (small active sets) -------------------------------------------
// The derivatives the backward pass needs at (x_t, u_t) -- A, B, l_x, l_u, l_xx, l_uu, l_ux, and V_x, V_xx at x_T
// (ilqr.hpp:92-100,106-113) -- depend on the current trajectory only, not on the value function, so they are
// independent across time steps.  `linearize_point` computes the derivative block of ONE (problem, time step) pair,
// optionally split over G cooperating threads (FD stencil tasks g, g + G, ...; the analytic callbacks on g == 0), and
// hands every entry to `put(offset, value)`; the engine runs it for all (problem, t) pairs of a small active set at
// once (linearize_kernel) and stores the blocks in HBM.  `riccati_sweep_thread` is then the sequential part alone:
// per step it fetches the block with `get(t, offset)` and runs the same riccati_step as the fused kernels.  Every
// entry is produced by the very function the fused pass uses, so the results are bit-identical (tests: host emulation
// and GPU).  t == T addresses the terminal block [V_x | V_xx].
template <class M>
MAS_HD void deriv_task_info(unsigned mask, int task, int* off, int* cnt, bool* analytic) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  if (task < D::tB) {
    *off = D::oA + (task - D::tA) * NX;
    *cnt = NX;
    *analytic = (mask & D_A) != 0;
  } else if (task < D::tlx) {
    *off = D::oB + (task - D::tB) * NX;
    *cnt = NX;
    *analytic = (mask & D_B) != 0;
  } else if (task < D::tlu) {
    *off = D::olx + (task - D::tlx);
    *cnt = 1;
    *analytic = (mask & D_LX) != 0;
  } else if (task < D::tlxx) {
    *off = D::olu + (task - D::tlu);
    *cnt = 1;
    *analytic = (mask & D_LU) != 0;
  } else if (task < D::tluu) {
    *off = D::olxx + (task - D::tlxx);
    *cnt = 1;
    *analytic = (mask & D_LXX) != 0;
  } else if (task < D::tlux) {
    *off = D::oluu + (task - D::tluu);
    *cnt = 1;
    *analytic = (mask & D_LUU) != 0;
  } else {
    *off = D::olux + (task - D::tlux);
    *cnt = 1;
    *analytic = (mask & D_LUX) != 0;
  }
  (void)NU;
}

template <class M, class Put>
MAS_HD void linearize_point(const BatchView<M::NX, M::NU>& v, int p, int t, unsigned mask, int g, int G, const Put& put) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  double x[NX], u[NU];
#pragma unroll
  for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
  if (t == v.T) {  // terminal value (ilqr.hpp:92-100); the aliased symmetrisation (:102) belongs to the sweep
    double blk[D::n_terminal_tasks];
    for (int task = g; task < D::n_terminal_tasks; task += G) {
      const bool analytic = task < NX ? (mask & D_VX) != 0 : (mask & D_VXX) != 0;
      if (analytic) continue;
      fd_terminal_task<M>(mask, task, x, prm, blk);
      put(task, blk[task]);
    }
    if (g == 0) {
      if (mask & D_VX) {
        double vx[NX];
        M::v_x(x, prm, vx);
        for (int i = 0; i < NX; ++i) put(i, vx[i]);
      }
      if (mask & D_VXX) {
        double vxx[NX * NX];
        M::v_xx(x, prm, vxx);
        for (int i = 0; i < NX * NX; ++i) put(NX + i, vxx[i]);
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
  double blk[D::size];
  for (int task = g; task < D::n_tasks; task += G) {
    int off, cnt;
    bool analytic;
    deriv_task_info<M>(mask, task, &off, &cnt, &analytic);
    if (analytic) continue;
    fd_stage_task<M>(mask, task, x, u, t, prm, blk);
    for (int k = 0; k < cnt; ++k) put(off + k, blk[off + k]);
  }
  if (g == 0) {  // analytic callbacks (:106-113)
    if (mask & D_A) {
      double A[NX * NX];
      M::jac_x(x, u, prm, A);
      for (int i = 0; i < NX * NX; ++i) put(D::oA + i, A[i]);
    }
    if (mask & D_B) {
      double B[NX * NU];
      M::jac_u(x, u, prm, B);
      for (int i = 0; i < NX * NU; ++i) put(D::oB + i, B[i]);
    }
    if (mask & D_LX) {
      double g1[NX];
      M::l_x(x, u, t, prm, g1);
      for (int i = 0; i < NX; ++i) put(D::olx + i, g1[i]);
    }
    if (mask & D_LU) {
      double g2[NU];
      M::l_u(x, u, t, prm, g2);
      for (int i = 0; i < NU; ++i) put(D::olu + i, g2[i]);
    }
    if (mask & D_LXX) {
      double H[NX * NX];
      M::l_xx(x, u, t, prm, H);
      for (int i = 0; i < NX * NX; ++i) put(D::olxx + i, H[i]);
    }
    if (mask & D_LUU) {
      double H[NU * NU];
      M::l_uu(x, u, t, prm, H);
      for (int i = 0; i < NU * NU; ++i) put(D::oluu + i, H[i]);
    }
    if (mask & D_LUX) {
      double H[NU * NX];
      M::l_ux(x, u, t, prm, H);
      for (int i = 0; i < NU * NX; ++i) put(D::olux + i, H[i]);
    }
  }
}

// The sequential half: terminal value, then T Riccati steps on stored derivative blocks.  `fetch(t, blk)` fills blk
// (DerivBlock<M>::size doubles) with the block of step t (t == T: [V_x | V_xx]) -- the device stages the next block in
// shared memory while the current step computes.  Returns the regularisation retries.
template <class M, int MASK_CT, class Fetch>
MAS_HD int riccati_sweep_thread(const BatchView<M::NX, M::NU>& v, int p, const Fetch& fetch) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  const int T = v.T;
  int retries = 0;
  const double al_rho = HasConstraints<M>::value ? v.penalty[p] : 0.0;
  double blk[D::size];
  double v_x[NX], v_xx[NX * NX];
  fetch(T, blk);
#pragma unroll
  for (int i = 0; i < NX; ++i) v_x[i] = blk[i];
#pragma unroll
  for (int i = 0; i < NX * NX; ++i) v_xx[i] = blk[NX + i];
  symmetrize_aliased<NX>(v_xx);
  for (int t = T - 1; t >= 0; --t) {
    double x[NX], u[NU];
    if (HasConstraints<M>::value) {  // only the constraint terms of riccati_step look at (x_t, u_t)
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    } else {
#pragma unroll
      for (int i = 0; i < NX; ++i) x[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NU; ++i) u[i] = 0.0;
    }
    fetch(t, blk);
    retries += riccati_step<M, MASK_CT>(v, p, t, x, u, prm, al_rho, blk + D::oA, blk + D::oB, blk + D::olx, blk + D::olu, blk + D::olxx, blk + D::oluu,
                                        blk + D::olux, v_x, v_xx);
  }
  return retries;
}

// ---- the Riccati recursion over the lanes of a problem (column-parallel) ---------------------------------------------
// What is left of a backward step once the derivatives are precomputed is ~1,200 dependent instructions of one thread
// (2.2 us per step on B200, measured), and small active sets pay exactly that latency T times per iteration.
// RiccatiLanes spreads one problem's step over LG = NX lanes (lane j owns column j of V_xx, Q_xx, Q_ux, K and of the
// value update; small quantities -- q_x, q_u, Q_uu, V_x -- are computed redundantly by every lane; the columns of
// Q_uu^-1 are dealt out): every output element is produced by the same function, with the same operand order, as in
// riccati_step, so the results are bit-identical (host emulation + GPU tests).  Four exchanges per step go through a
// small per-problem area `xch` (shared memory on the device) with a group barrier after each phase.  Models with path
// constraints keep the one-thread step (the augmented-Lagrangian terms touch every block of Q).
template <int N>
MAS_HD void llt_inverse_col(const double* L, int c, double* x) {  // column c of llt_inverse<N>
#pragma unroll
  for (int i = 0; i < N; ++i) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double s = x[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= L[i + j * N] * x[j];
    x[i] = pm::div_(s, L[i + i * N]);
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int j = i + 1; j < N; ++j) s -= L[j + i * N] * x[j];
    x[i] = pm::div_(s, L[i + i * N]);
  }
}
// column j (run-time) of mat_nn_sb<R, KD, CC, NZ>(A, B, C): out[i] = sum over the k with bit (k + j*KD) of NZ set of
// A[i + k*R] * bcol[k], same order, same "first term starts the sum" rule; no branches on j
template <int R, int KD>
MAS_HD void mat_nn_sb_col(const double* A, const double* bcol, unsigned long long nz, int j, double* out) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    double s = 0.0;
    bool first = true;
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const bool on = (nz >> (k + j * KD)) & 1ull;
      const double prod = A[i + k * R] * bcol[k];
      const double sum = s + prod;
      s = on ? (first ? prod : sum) : s;
      first = first && !on;
    }
    out[i] = s;
  }
}

template <class M, int MASK_CT>
struct RiccatiLanes {
  static constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  static constexpr int LG = NX <= 1 ? 1 : (NX <= 2 ? 2 : (NX <= 4 ? 4 : 8));  // lanes per problem (power of two >= NX)
  static_assert(NX <= 8, "RiccatiLanes: one lane per state column, up to 8");
  // exchange area (doubles): A^T V | B^T V | inv | K | Q_ux | V_xx
  static constexpr int xAtV = 0, xBtV = xAtV + NX * NX, xInv = xBtV + NU * NX, xK = xInv + NU * NU, xQux = xK + NU * NX, xV = xQux + NU * NX,
                       XCH = xV + NX * NX;
  static constexpr unsigned long long kDense = ~0ull;
  static constexpr unsigned long long a_nz = (MASK_CT >= 0 && (MASK_CT & D_A)) ? M::A_NZ : kDense;
  static constexpr unsigned long long b_nz = (MASK_CT >= 0 && (MASK_CT & D_B)) ? M::B_NZ : kDense;

  // per-lane state
  double v_x[NX], vcol[NX];
  double q_x[NX], q_u[NU], q_uu[NU * NU], L[NU * NU], qxx_col[NX], qux_col[NU], k_col[NU], kv[NU];
  int retries = 0;

  // terminal block [V_x | V_xx]: every lane symmetrises the whole matrix and keeps its column (ilqr.hpp:92-102)
  MAS_HD void init_terminal(const double* blk, int j) {
    double vxx[NX * NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) v_x[i] = blk[i];
#pragma unroll
    for (int i = 0; i < NX * NX; ++i) vxx[i] = blk[NX + i];
    symmetrize_aliased<NX>(vxx);
#pragma unroll
    for (int i = 0; i < NX; ++i) vcol[i] = 0.0;
#pragma unroll
    for (int c = 0; c < NX; ++c) {
      if (c == j) {
#pragma unroll
        for (int i = 0; i < NX; ++i) vcol[i] = vxx[i + c * NX];
      }
    }
  }
  // (A^T V)(:, j), (B^T V)(:, j)                                                      ilqr.hpp:117-119, first factors
  MAS_HD void phase_a(const double* blk, int j, double* xch) const {
    double atv[NX], btv[NU];
    mat_tn_sa<NX, NX, 1, a_nz>(blk + D::oA, vcol, atv);
    mat_tn_sa<NX, NU, 1, b_nz>(blk + D::oB, vcol, btv);
#pragma unroll
    for (int i = 0; i < NX; ++i) xch[xAtV + i + j * NX] = atv[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) xch[xBtV + i + j * NU] = btv[i];
  }
  // Q blocks (:115-119), Q_uu regularisation + LLT (:172-183), my columns of the inverse
  MAS_HD void phase_b(const double* blk, int j, double* xch) {
    const double* A = blk + D::oA;
    const double* B = blk + D::oB;
    double tmp[NX > NU ? NX : NU];
    mat_tn_sa<NX, NX, 1, a_nz>(A, v_x, tmp);
#pragma unroll
    for (int i = 0; i < NX; ++i) q_x[i] = blk[D::olx + i] + tmp[i];
    mat_tn_sa<NX, NU, 1, b_nz>(B, v_x, tmp);
#pragma unroll
    for (int i = 0; i < NU; ++i) q_u[i] = blk[D::olu + i] + tmp[i];
    double acol[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k) acol[k] = A[k + j * NX];
    double t4[NX], t2[NU];
    mat_nn_sb_col<NX, NX>(xch + xAtV, acol, a_nz, j, t4);
#pragma unroll
    for (int i = 0; i < NX; ++i) qxx_col[i] = blk[D::olxx + i + j * NX] + t4[i];
    mat_nn_sb_col<NU, NX>(xch + xBtV, acol, a_nz, j, t2);
#pragma unroll
    for (int i = 0; i < NU; ++i) qux_col[i] = blk[D::olux + i + j * NU] + t2[i];
    double tuu[NU * NU];
    mat_nn_sb<NU, NX, NU, b_nz>(xch + xBtV, B, tuu);
#pragma unroll
    for (int i = 0; i < NU * NU; ++i) q_uu[i] = blk[D::oluu + i] + tuu[i];
    double q_reg[NU * NU];
#pragma unroll
    for (int i = 0; i < NU * NU; ++i) q_reg[i] = q_uu[i];
    double reg = 1e-6;
    while (!llt_factor<NU>(q_reg, L)) {
#pragma unroll
      for (int i = 0; i < NU; ++i) q_reg[i + i * NU] += reg;
      reg *= 10.0;
      if (j == 0) ++retries;
      if (!(reg < 1e300)) break;
    }
    for (int c = j; c < NU; c += LG) {
      double x[NU];
      llt_inverse_col<NU>(L, c, x);
#pragma unroll
      for (int i = 0; i < NU; ++i) xch[xInv + i + c * NU] = x[i];
    }
  }
  // gains (:185-186): k on every lane, my column of K
  MAS_HD void phase_c(int j, double* xch) {
    double ninv[NU * NU];
#pragma unroll
    for (int i = 0; i < NU * NU; ++i) ninv[i] = -xch[xInv + i];
    mat_nn<NU, NU, 1>(ninv, q_u, kv);
    mat_nn<NU, NU, 1>(ninv, qux_col, k_col);
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      xch[xK + i + j * NU] = k_col[i];
      xch[xQux + i + j * NU] = qux_col[i];
    }
  }
  // gains to HBM, value update (:188-191): V_x on every lane, my column of the unsymmetrised V_xx
  MAS_HD void phase_d(const BatchView<NX, NU>& v, int p, int t, int j, double* xch) {
    const double* Km = xch + xK;
    const double* q_ux = xch + xQux;
    if (j == 0) {
#pragma unroll
      for (int i = 0; i < NU; ++i) v.kff[soa_index<NU>(t, i, v.ld, p)] = kv[i];
    }
#pragma unroll
    for (int i = 0; i < NU; ++i) v.K[soa_index<NU * NX>(t, i + j * NU, v.ld, p)] = k_col[i];
    double KtQuu[NX * NU], t1[NX], t2[NX], t3[NX];
    mat_tn<NU, NX, NU>(Km, q_uu, KtQuu);
    mat_tn<NU, NX, 1>(Km, q_u, t1);
    mat_tn<NU, NX, 1>(q_ux, kv, t2);
    mat_nn<NX, NU, 1>(KtQuu, kv, t3);
#pragma unroll
    for (int i = 0; i < NX; ++i) v_x[i] = ((q_x[i] + t1[i]) + t2[i]) + t3[i];
    double m1[NX], m2[NX], m3[NX];
    mat_tn<NU, NX, 1>(Km, qux_col, m1);
    mat_tn<NU, NX, 1>(q_ux, k_col, m2);
    mat_nn<NX, NU, 1>(KtQuu, k_col, m3);
#pragma unroll
    for (int i = 0; i < NX; ++i) xch[xV + i + j * NX] = ((qxx_col[i] + m1[i]) + m2[i]) + m3[i];
  }
  // `v_xx = 0.5 * (v_xx + v_xx.transpose())` evaluated in place, columns outer (symmetrize_aliased), column j of the result:
  // below the diagonal both operands are old; above it the transposed operand is the already-updated lower entry
  MAS_HD void phase_e(int j, const double* xch) {
    const double* V = xch + xV;
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const double a = V[i + j * NX], b = V[j + i * NX];
      const double lower = 0.5 * (b + a);  // the value entry (j, i) took when column i < j was processed
      vcol[i] = (i < j) ? 0.5 * (a + lower) : 0.5 * (a + b);
    }
  }
};

// ---- the Riccati recursion over NX*NX lanes of a problem (element-parallel) ------------------------------------------
// RiccatiLanes leaves every lane ~340 dependent fp64 instructions per step (q_x, q_u, Q_uu, the factorisation and K^T Q_uu
// are computed redundantly), and a lone warp issues them at ~5.5 cycles each: 2.1 us per step, T times per iteration,
// whatever the problem count.  RiccatiWide gives a problem one lane per ENTRY (i, j) of the NX x NX matrices: a lane
// produces one entry of A^T V_xx, Q_xx, m1, m2, m3 and V_xx per step plus one entry of the smaller products (B^T V_xx,
// Q_ux, Q_uu, K, K^T Q_uu, q_x, q_u, k, V_x), dealt out by lane index; only the NU x NU factorisation + inverse is
// repeated on every lane.  Every entry is the same k-ascending sum with the same structural zeros left out as in
// riccati_step (sparse_dot == one entry of mat_tn_sa / mat_nn_sb / mat_tn / mat_nn), so the results are bit-identical
// (host emulation + GPU tests).  Operands travel through a per-problem area of shared memory, one barrier per phase.
// a[k*sa] * b[k*sb] summed over the k whose bit (bit0 + k) of nz is set, k ascending, the first term starts the sum;
// branch-free so that the lanes of a warp may differ in (a, b, nz, bit0)
template <int KD>
MAS_HD double sparse_dot(const double* a, int sa, const double* b, int sb, unsigned long long nz, int bit0) {
  double s = 0.0;
  bool first = true;
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const bool on = (nz >> (bit0 + k)) & 1ull;
    const double prod = a[k * sa] * b[k * sb];
    const double sum = s + prod;
    s = on ? (first ? prod : sum) : s;
    first = first && !on;
  }
  return s;
}
template <int KD>
MAS_HD double dense_dot(const double* a, int sa, const double* b, int sb) {
  double s = a[0] * b[0];
#pragma unroll
  for (int k = 1; k < KD; ++k) s = s + a[k * sa] * b[k * sb];
  return s;
}

template <class M, int MASK_CT>
struct RiccatiWide {
  static constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  static constexpr int LW = NX * NX;  // lanes per problem
  static constexpr bool kSupported = (LW == 4 || LW == 8 || LW == 16 || LW == 32) && NU * NX + NX + NU <= LW && NU * NX + NU * NU <= LW &&
                                     NU * NX + 2 * NX <= LW && !HasConstraints<M>::value;
  // lane roles beside the (i, j) entry every lane owns: [0, UX) an entry of the NU x NX (or NX x NU) matrices, then ranges of NX / NU lanes
  static constexpr int UX = NU * NX, rQx = UX, rQu = UX + NX, rQuu = UX, rKv = UX, rT1 = UX, rT2 = UX + NX;
  // exchange area (doubles)
  static constexpr int xAtV = 0, xBtV = xAtV + NX * NX, xQux = xBtV + UX, xQuu = xQux + UX, xqu = xQuu + NU * NU, xK = xqu + NU, xkv = xK + UX,
                       xKtQuu = xkv + NU, xt2 = xKtQuu + UX, xVu = xt2 + NX, xV = xVu + NX * NX, xvx = xV + NX * NX, XCH = xvx + NX;
  static constexpr unsigned long long kDense = ~0ull;
  static constexpr unsigned long long a_nz = (MASK_CT >= 0 && (MASK_CT & D_A)) ? M::A_NZ : kDense;
  static constexpr unsigned long long b_nz = (MASK_CT >= 0 && (MASK_CT & D_B)) ? M::B_NZ : kDense;

  // per-lane state
  double qxx = 0.0;   // Q_xx(i, j), then (Q_xx + m1) + m2
  double qx = 0.0;    // lanes [rQx, rQx + NX): q_x(i), then q_x + t1
  int retries = 0;

  // terminal block [V_x | V_xx] (ilqr.hpp:92-100); the aliased symmetrisation (:102) is phase 6
  static MAS_HD void init_terminal(const double* blk, int e, double* xch) {
    xch[xVu + e] = blk[NX + e];
    if (e < NX) xch[xvx + e] = blk[e];
  }
  // first factors (:115-119): (A^T V_xx)(i, j); (B^T V_xx)(r, j); q_x(i) = l_x + A^T V_x; q_u(r) = l_u + B^T V_x
  MAS_HD void phase1(const double* blk, int e, double* xch) {
    const int i = e % NX, j = e / NX;
    xch[xAtV + e] = sparse_dot<NX>(blk + D::oA + i * NX, 1, xch + xV + j * NX, 1, a_nz, i * NX);
    if (e < rQu + NU) {
      const bool isBtV = e < UX, isQx = !isBtV && e < rQu;
      const int r = isBtV ? e % NU : e - rQu, jb = e / NU, ix = e - rQx;
      const int col = isQx ? ix : r;  // column of A (q_x) or of B (B^T V, q_u)
      const double* a = blk + (isQx ? D::oA : D::oB) + col * NX;
      const double* b = isBtV ? xch + xV + jb * NX : xch + xvx;
      const double dot = sparse_dot<NX>(a, 1, b, 1, isQx ? a_nz : b_nz, col * NX);
      if (isBtV) {
        xch[xBtV + e] = dot;
      } else if (isQx) {
        qx = blk[D::olx + ix] + dot;
      } else {
        xch[xqu + r] = blk[D::olu + r] + dot;
      }
    }
  }
  // second factors (:117-119): Q_xx(i, j); Q_ux(r, j); Q_uu(r, s)
  MAS_HD void phase2(const double* blk, int e, double* xch) {
    const int i = e % NX, j = e / NX;
    qxx = blk[D::olxx + e] + sparse_dot<NX>(xch + xAtV + i, NX, blk + D::oA + j * NX, 1, a_nz, j * NX);
    if (e < rQuu + NU * NU) {
      const bool isQux = e < UX;
      const int f = isQux ? e : e - rQuu, r = f % NU, c = f / NU;  // c: column of A (Q_ux) or of B (Q_uu)
      const double dot = sparse_dot<NX>(xch + xBtV + r, NU, blk + (isQux ? D::oA : D::oB) + c * NX, 1, isQux ? a_nz : b_nz, c * NX);
      if (isQux) xch[xQux + f] = blk[D::olux + f] + dot;
      else xch[xQuu + f] = blk[D::oluu + f] + dot;
    }
  }
  // Q_uu regularisation + LLT + inverse on every lane (:172-183), then the gains (:185-186): K(r, j), k(r)
  MAS_HD void phase3(const BatchView<NX, NU>& v, int p, int t, int e, double* xch) {
    double q_reg[NU * NU], L[NU * NU], inv[NU * NU];
#pragma unroll
    for (int n = 0; n < NU * NU; ++n) q_reg[n] = xch[xQuu + n];
    double reg = 1e-6;
    while (!llt_factor<NU>(q_reg, L)) {
#pragma unroll
      for (int n = 0; n < NU; ++n) q_reg[n + n * NU] += reg;
      reg *= 10.0;
      if (e == 0) ++retries;
      if (!(reg < 1e300)) break;
    }
    llt_inverse<NU>(L, inv);
    if (e < rKv + NU) {
      const bool isK = e < UX;
      const int r = isK ? e % NU : e - rKv, j = e / NU;
      double nrow[NU];  // row r of -Q_uu^-1, picked with selects (r is a run-time value)
#pragma unroll
      for (int sidx = 0; sidx < NU; ++sidx) {
        double val = -inv[0 + sidx * NU];
#pragma unroll
        for (int rr = 1; rr < NU; ++rr) val = (r == rr) ? -inv[rr + sidx * NU] : val;
        nrow[sidx] = val;
      }
      const double g = dense_dot<NU>(nrow, 1, isK ? xch + xQux + j * NU : xch + xqu, 1);
      if (isK) {
        xch[xK + e] = g;
        v.K[soa_index<NU * NX>(t, e, v.ld, p)] = g;
      } else {
        xch[xkv + r] = g;
        v.kff[soa_index<NU>(t, r, v.ld, p)] = g;
      }
    }
  }
  // value update, first half (:188-191): m1, m2 of my entry; (K^T Q_uu)(i, s); K^T q_u; Q_ux^T k
  MAS_HD void phase4(int e, double* xch) {
    const int i = e % NX, j = e / NX;
    const double m1 = dense_dot<NU>(xch + xK + i * NU, 1, xch + xQux + j * NU, 1);
    const double m2 = dense_dot<NU>(xch + xQux + i * NU, 1, xch + xK + j * NU, 1);
    qxx = (qxx + m1) + m2;
    if (e < rT2 + NX) {
      const bool isKtQ = e < UX, isT1 = !isKtQ && e < rT2;
      const int ii = isKtQ ? e % NX : (isT1 ? e - rT1 : e - rT2), sc = e / NX;
      const double* a = (isKtQ || isT1) ? xch + xK + ii * NU : xch + xQux + ii * NU;
      const double* b = isKtQ ? xch + xQuu + sc * NU : (isT1 ? xch + xqu : xch + xkv);
      const double d = dense_dot<NU>(a, 1, b, 1);
      if (isKtQ) xch[xKtQuu + e] = d;
      else if (isT1) qx = qx + d;
      else xch[xt2 + ii] = d;
    }
  }
  // value update, second half: m3 and the unsymmetrised V_xx(i, j); V_x(i)
  MAS_HD void phase5(int e, double* xch) {
    const int i = e % NX, j = e / NX;
    const double m3 = dense_dot<NU>(xch + xKtQuu + i, NX, xch + xK + j * NU, 1);
    xch[xVu + e] = qxx + m3;
    if (e >= rT1 && e < rT1 + NX) {
      const int ii = e - rT1;
      const double t3 = dense_dot<NU>(xch + xKtQuu + ii, NX, xch + xkv, 1);
      xch[xvx + ii] = (qx + xch[xt2 + ii]) + t3;
    }
  }
  // `v_xx = 0.5 * (v_xx + v_xx.transpose())` evaluated in place, columns outer (symmetrize_aliased), entry (i, j): below the
  // diagonal both operands are old; above it the transposed operand is the already-updated lower entry
  static MAS_HD void phase6(int e, double* xch) {
    const int i = e % NX, j = e / NX;
    const double a = xch[xVu + i + j * NX], b = xch[xVu + j + i * NX];
    const double lower = 0.5 * (b + a);
    xch[xV + e] = (i < j) ? 0.5 * (a + lower) : 0.5 * (a + b);
  }
};

// ---- forward pass (ilqr.hpp:206-217) for C step sizes at once, merit only ------------------------
// The C rollouts share the loads of the nominal trajectory and gains and give the fp64 pipe C
// independent dependency chains.  merit[c] = sum_t stage + terminal, accumulated in t order.
// STORE: chain c also writes its controls and states to trial slot slot0 + c * slot_stride (BatchView::trial_*).
// objective (optional) receives the plain cost of each chain (== merit for models without constraints).
template <class M, int C, bool STORE = false>
MAS_HD void trial_rollout(const BatchView<M::NX, M::NU>& v, int p, const double* prm, const double* alpha, double* merit, int slot0 = -1,
                          int slot_stride = 0, double* objective = nullptr, double* stage = nullptr) {
  constexpr int NX = M::NX, NU = M::NU;
  constexpr bool kAL = HasConstraints<M>::value;
  const size_t n_slots = static_cast<size_t>(v.trial_slots);
  if (stage) stage_issue_step<NX, NU>(v, p, 0, 0, stage, 0);
  double xt[C][NX], cost[C];
  double al_terms[kAL ? C : 1][kAL ? 3 * kMaxALHorizon : 1];  // merit addends per step (local memory, constrained models only)
  const double al_rho = kAL ? v.penalty[p] : 0.0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    cost[c] = 0.0;
#pragma unroll
    for (int i = 0; i < NX; ++i) xt[c][i] = v.x0[static_cast<size_t>(i) * v.ld + p];
  }
  const size_t ld = static_cast<size_t>(v.ld);
  const double *rX = v.X + p, *rU = v.U + p, *rk = v.kff + p, *rK = v.K + p;  // rows of step t (plain-load path)
  for (int t = 0; t < v.T; ++t) {
    double xn[NX], un[NU], kv[NU], Km[NU * NX];
    if (stage) {
      stage_wait();
      stage_read_step<NX, NU>(stage, t & 1, xn, un, kv, Km);
      if (t + 1 < v.T) stage_issue_step<NX, NU>(v, p, t + 1, t + 1, stage, (t + 1) & 1);
    } else {
      // running row pointers (one add per array and step instead of a 64-bit index product per element); the rows of
      // step t+1 are prefetched without a branch: on the last step the offset is zero and the current rows are named again
#pragma unroll
      for (int i = 0; i < NX; ++i) xn[i] = rX[i * ld];
#pragma unroll
      for (int i = 0; i < NU; ++i) un[i] = rU[i * ld];
#pragma unroll
      for (int i = 0; i < NU; ++i) kv[i] = rk[i * ld];
#pragma unroll
      for (int i = 0; i < NU * NX; ++i) Km[i] = rK[i * ld];
      const size_t nxt = t + 1 < v.T ? 1 : 0;
      const double *nX = rX + nxt * (NX * ld), *nU = rU + nxt * (NU * ld), *nk = rk + nxt * (NU * ld), *nK = rK + nxt * (NU * NX * ld);
#pragma unroll
      for (int i = 0; i < NX; ++i) prefetch_l1(&nX[i * ld]);
#pragma unroll
      for (int i = 0; i < NU; ++i) prefetch_l1(&nU[i * ld]);
#pragma unroll
      for (int i = 0; i < NU; ++i) prefetch_l1(&nk[i * ld]);
#pragma unroll
      for (int i = 0; i < NU * NX; ++i) prefetch_l1(&nK[i * ld]);
      rX = nX;
      rU = nU;
      rk = nk;
      rK = nK;
    }
    double u[C][NU], xnext[C][NX];
    bool exact = true;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double dx[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) dx[i] = xt[c][i] - xn[i];
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        double kdx = Km[i + 0 * NU] * dx[0];
#pragma unroll
        for (int j = 1; j < NX; ++j) kdx = kdx + Km[i + j * NU] * dx[j];
        double ui = (un[i] + alpha[c] * kv[i]) + kdx;
        ui = (v.clamp_hi[i] < ui) ? v.clamp_hi[i] : ui;  // clamp_controls: cwiseMin(upper) then cwiseMax(lower); +-inf = no bounds
        ui = (v.clamp_lo[i] > ui) ? v.clamp_lo[i] : ui;
        u[c][i] = ui;
      }
      cost[c] += M::stage(xt[c], u[c], t, prm);
      if (kAL) al_merit_addends<M>(v, p, t, xt[c], u[c], prm, al_rho, &al_terms[c][3 * t]);
      const bool e = rk4_step_spec<M>(xt[c], u[c], prm, v.dt, v.half_dt, v.sixth_dt, xnext[c]);  // one basic block for all C candidates
      exact = exact && e;
    }
    if (!exact) {  // a division off its fast path (non-finite trajectories): the step again, branch by branch
#pragma unroll
      for (int c = 0; c < C; ++c) rk4_step<M>(xt[c], u[c], prm, v.dt, xnext[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int i = 0; i < NX; ++i) xt[c][i] = xnext[c][i];
      if (STORE) {
        const size_t slot = static_cast<size_t>(slot0 + c * slot_stride);
#pragma unroll
        for (int i = 0; i < NU; ++i) store_streaming(&v.trial_U[(static_cast<size_t>(t) * NU + i) * n_slots + slot], u[c][i]);
#pragma unroll
        for (int i = 0; i < NX; ++i) store_streaming(&v.trial_X[(static_cast<size_t>(t) * NX + i) * n_slots + slot], xnext[c][i]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    cost[c] += M::terminal(xt[c], prm);
    if (objective) objective[c] = cost[c];
    merit[c] = kAL ? al_finish_merit<M>(cost[c], al_terms[c], v.T) : cost[c];
  }
}

// Accepted step taken from a trial slot: U[t] and X[t+1] of problem p are overwritten with the stored candidate.
template <class M>
MAS_HD void accept_stored(const BatchView<M::NX, M::NU>& v, int p, int slot) {
  constexpr int NX = M::NX, NU = M::NU, kChunk = 8;
  const size_t n_slots = static_cast<size_t>(v.trial_slots), s = static_cast<size_t>(slot);
  // all loads of a chunk of steps are issued before its first store: the pointers may alias as far as the compiler
  // knows, and a store waiting for its operand would hold back every load behind it (one DRAM latency per element)
  for (int t0 = 0; t0 < v.T; t0 += kChunk) {
    double bu[kChunk][NU], bx[kChunk][NX];
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      const int t = t0 + k < v.T ? t0 + k : v.T - 1;
#pragma unroll
      for (int i = 0; i < NU; ++i) bu[k][i] = load_streaming(&v.trial_U[(static_cast<size_t>(t) * NU + i) * n_slots + s]);
#pragma unroll
      for (int i = 0; i < NX; ++i) bx[k][i] = load_streaming(&v.trial_X[(static_cast<size_t>(t) * NX + i) * n_slots + s]);
    }
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      const int t = t0 + k;
      if (t < v.T) {
#pragma unroll
        for (int i = 0; i < NU; ++i) v.U[soa_index<NU>(t, i, v.ld, p)] = bu[k][i];
#pragma unroll
        for (int i = 0; i < NX; ++i) v.X[soa_index<NX>(t + 1, i, v.ld, p)] = bx[k][i];
      }
    }
  }
}

// ---- accepted step: same arithmetic as trial_rollout, writing X and U in place -------------------
// Step t reads the nominal x_t, u_t, K_t, k_t before U[t] and X[t+1] are overwritten, and the
// nominal x_{t+1} is fetched before X[t+1] is stored, so one buffer serves as old and new trajectory.
template <class M>
MAS_HD double commit_rollout(const BatchView<M::NX, M::NU>& v, int p, const double* prm, double alpha) {
  constexpr int NX = M::NX, NU = M::NU;
  double xt[NX], xn[NX], cost = 0.0;
#pragma unroll
  for (int i = 0; i < NX; ++i) {
    xt[i] = v.x0[static_cast<size_t>(i) * v.ld + p];
    xn[i] = v.X[soa_index<NX>(0, i, v.ld, p)];
  }
  for (int t = 0; t < v.T; ++t) {
    double un[NU], kv[NU], Km[NU * NX], xn_next[NX];
#pragma unroll
    for (int i = 0; i < NU; ++i) un[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NU; ++i) kv[i] = v.kff[soa_index<NU>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NU * NX; ++i) Km[i] = v.K[soa_index<NU * NX>(t, i, v.ld, p)];
#pragma unroll
    for (int i = 0; i < NX; ++i) xn_next[i] = v.X[soa_index<NX>(t + 1, i, v.ld, p)];
    if (t + 1 < v.T) {
#pragma unroll
      for (int i = 0; i < NX; ++i) prefetch_l1(&v.X[soa_index<NX>(t + 2, i, v.ld, p)]);
#pragma unroll
      for (int i = 0; i < NU; ++i) prefetch_l1(&v.U[soa_index<NU>(t + 1, i, v.ld, p)]);
#pragma unroll
      for (int i = 0; i < NU; ++i) prefetch_l1(&v.kff[soa_index<NU>(t + 1, i, v.ld, p)]);
#pragma unroll
      for (int i = 0; i < NU * NX; ++i) prefetch_l1(&v.K[soa_index<NU * NX>(t + 1, i, v.ld, p)]);
    }
    double dx[NX], u[NU], xnext[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) dx[i] = xt[i] - xn[i];
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      double kdx = Km[i + 0 * NU] * dx[0];
#pragma unroll
      for (int j = 1; j < NX; ++j) kdx = kdx + Km[i + j * NU] * dx[j];
      double ui = (un[i] + alpha * kv[i]) + kdx;
      ui = (v.clamp_hi[i] < ui) ? v.clamp_hi[i] : ui;
      ui = (v.clamp_lo[i] > ui) ? v.clamp_lo[i] : ui;
      u[i] = ui;
      v.U[soa_index<NU>(t, i, v.ld, p)] = ui;
    }
    cost += M::stage(xt, u, t, prm);
    if (!rk4_step_spec<M>(xt, u, prm, v.dt, v.half_dt, v.sixth_dt, xnext)) rk4_step<M>(xt, u, prm, v.dt, xnext);
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      xt[i] = xnext[i];
      xn[i] = xn_next[i];
      v.X[soa_index<NX>(t + 1, i, v.ld, p)] = xnext[i];
    }
  }
  cost += M::terminal(xt, prm);
  return cost;
}

// Step size of candidate j: alpha starts at 1.0 and is halved (ilqr.hpp:200,227) -> exact 2^-j.
MAS_HD double alpha_of(int j) {
  double a = 1.0;
  for (int i = 0; i < j; ++i) a *= 0.5;
  return a;
}

// One lane's share of the line search: candidates j = lane, lane + L, ... in chunks of C, stopping
// after the first chunk that holds an improving candidate (later candidates of this lane cannot win).
// Returns the lane's first improving candidate (kNumAlphas if none) and its merit.
// slot0 >= 0: the lane's chains write their trajectories to trial slots slot0 + c * slot_stride (each pass overwrites
// the previous one, which held no improving candidate); *best_slot / *best_objective describe the winner.
template <class M, int L, int C, bool STORE = false>
MAS_HD void lane_line_search(const BatchView<M::NX, M::NU>& v, int p, const double* prm, int lane, double current_merit, int* best_j,
                             double* best_merit, int slot0 = -1, int slot_stride = 0, int* best_slot = nullptr, double* best_objective = nullptr,
                             double* stage = nullptr) {
  *best_j = kNumAlphas;
  *best_merit = current_merit;
  if (best_slot) *best_slot = -1;
  for (int base = lane; base < kNumAlphas; base += L * C) {
    double alpha[C], merit[C], objective[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = base + c * L;
      alpha[c] = alpha_of(j < kNumAlphas ? j : kNumAlphas - 1);
    }
    trial_rollout<M, C, STORE>(v, p, prm, alpha, merit, slot0, slot_stride, objective, stage);
    bool found = false;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = base + c * L;
      if (!found && j < kNumAlphas && merit[c] < current_merit) {  // strict <, NaN never improves (:220)
        found = true;
        *best_j = j;
        *best_merit = merit[c];
        if (best_slot) *best_slot = STORE ? slot0 + c * slot_stride : -1;
        if (best_objective) *best_objective = objective[c];
      }
    }
    if (found) break;
  }
}

// ---- warp-cooperative line search: task assignment ----------------------------------------------------
// A warp owns up to 32 problems; a task is (problem, step-size index).  Each round the 32 lanes are dealt out
// evenly to the problems still searching: a problem's share q covers its next q untried step sizes, so a
// problem that needs many candidates gets them evaluated side by side by the lanes of problems that are
// already done.  Sequential semantics are preserved by the owner scanning its results in index order.
// Arrays are per warp (shared memory on the device); `next` = first untried index, `done` = search over.
struct CoopPlan {
  int owner[32];   // lane -> owning problem slot, -1 = idle this round
  int cand[32];    // lane -> step-size index
  int quota[32];   // slot -> candidates evaluated this round
};
// With `chains` > 1 a lane's task is `chains` consecutive step sizes of one problem, rolled out by one thread as
// independent instruction streams (they share the nominal trajectory and the gains it loads): quota stays in lanes.
MAS_HD void coop_assign(const int* done, const int* next, int n_valid, CoopPlan* plan, int chains = 1) {
  int n_act = 0;
  for (int i = 0; i < n_valid; ++i) n_act += done[i] ? 0 : 1;
  for (int l = 0; l < 32; ++l) {
    plan->owner[l] = -1;
    plan->cand[l] = 0;
    plan->quota[l] = 0;
  }
  if (n_act == 0) return;
  const int base = 32 / n_act, rem = 32 % n_act;
  int lane = 0, rank = 0;
  for (int i = 0; i < n_valid; ++i) {
    if (done[i]) continue;
    int q = base + (rank < rem ? 1 : 0);
    const int left = (kNumAlphas - next[i] + chains - 1) / chains;
    if (q > left) q = left;
    plan->quota[i] = q;
    for (int k = 0; k < q; ++k) {
      plan->owner[lane] = i;
      plan->cand[lane] = next[i] + k * chains;
      ++lane;
    }
    ++rank;
  }
}

// Owner's verdict after a round: first improving candidate among those just evaluated (strict <, ilqr.hpp:220).
// Returns true when the problem's search is over (accepted, or all ten tried); *accepted = index or -1.
MAS_HD bool coop_owner_update(const double* merits /* [kNumAlphas] of this problem */, double current_merit, int quota, int* next, int* accepted,
                              double* accepted_merit, int chains = 1) {
  int end = *next + quota * chains;
  if (end > kNumAlphas) end = kNumAlphas;
  for (int j = *next; j < end; ++j)
    if (merits[j] < current_merit) {
      *accepted = j;
      *accepted_merit = merits[j];
      return true;
    }
  *next = end;
  return end >= kNumAlphas;
}

// Accept / bookkeeping / stop test for one problem (ilqr.hpp:230-234,269-271).  Returns true when
// the problem needs another iteration.
// best_slot >= 0: the accepted candidate's trajectory is in that trial slot and its cost is best_objective;
// kSlotCommitted: the caller has already copied it into X, U; otherwise (-1) it is rolled out again.
constexpr int kSlotCommitted = -2;
template <class M>
MAS_HD bool finish_iteration(const BatchView<M::NX, M::NU>& v, int p, const double* prm, double current_merit, int best_j, double best_merit,
                             int best_slot = -1, double best_objective = 0.0) {
  if (best_j < kNumAlphas) {
    double c;
    if (best_slot >= 0 || best_slot == kSlotCommitted) {
      if (best_slot >= 0) accept_stored<M>(v, p, best_slot);
      c = best_objective;
    } else {
      c = commit_rollout<M>(v, p, prm, alpha_of(best_j));
    }
    v.cost[p] = c;  // objective(x,u) on the accepted trajectory: the trial's own sum, or the same arithmetic again
    v.merit[p] = best_merit;
  }
  const double improvement = current_merit - best_merit;
  const int it = v.iters[p] + 1;
  v.iters[p] = it;
  v.trials[p] += (best_j < kNumAlphas) ? best_j + 1 : kNumAlphas;
  bool feasible = true;
  double eq_norm = 0.0, ineq_norm = 0.0;
  if (HasConstraints<M>::value) {
    // multipliers <- multipliers + rho * residual on the new trajectory, penalty growth, and the violation
    // norms that gate the stop test (ilqr.hpp:236-260,269-270)
    al_update<M>(v, p, prm, &eq_norm, &ineq_norm);
    feasible = eq_norm < v.constraint_tolerance && ineq_norm < v.constraint_tolerance;
  }
  if (v.dbg && it < v.dbg_records) {  // what `debug` prints after the iteration (ilqr.hpp:262-267)
    double* r = v.dbg + static_cast<size_t>(it) * kDebugFields * v.ld + p;
    r[0 * static_cast<size_t>(v.ld)] = v.cost[p];
    r[1 * static_cast<size_t>(v.ld)] = best_merit;
    r[2 * static_cast<size_t>(v.ld)] = improvement;
    r[3 * static_cast<size_t>(v.ld)] = eq_norm;
    r[4 * static_cast<size_t>(v.ld)] = ineq_norm;
    r[5 * static_cast<size_t>(v.ld)] = (best_j < kNumAlphas) ? static_cast<double>(best_j) : -1.0;
  }
  if (improvement < v.tolerance && feasible) {
    v.status[p] = STATUS_CONVERGED;
    return false;
  }
  if (it >= v.max_iterations) {
    v.status[p] = STATUS_MAX_ITER;
    return false;
  }
  return true;
}

}  // namespace mas_b200
