// stacked_mixed.cuh -- CentralizedStrategy (strategies/centralized.hpp:18-38) for agents of DIFFERENT registered models.
//
// MultiAgentProblem::build_global_ocp (multi_agent_problem.hpp:52-127) stacks any mix of agents: blocks in id order
// (compute_offsets, :37-50), horizon and dt of the first block, bounds only when every agent has both, block-diagonal
// dynamics, stage / terminal costs summed in block order starting from 0.0, and -- because the stacked OCP installs none
// of the agents' analytic callbacks -- the finite-difference default for every derivative (ocp.hpp:117-135).  The reference
// then runs iLQR::solve (ilqr.hpp:59-273) on that OCP.  centralized.cuh does this for agents of ONE model, with the block
// structure compiled in; this file is the general case: block shapes and models are run-time data (MixedBlock), every
// model call goes through a switch on the model id, and all matrices are dense, i.e. every sum has exactly the terms, in
// exactly the order, of the reference's dense Eigen expressions (k ascending, the first product starts the sum).
//
// One CTA per scenario, data-parallel phases `for (idx = tid; idx < n; idx += nthr)` between barriers with all state in the
// scenario's workspace, so tests/csrc/host_emulation.cpp runs the same source with tid = 0, nthr = 1.
//
// Cost of a finite-difference point: a perturbation touches one or two blocks; the stacked value is re-formed as the same
// left-to-right sum over ALL blocks, untouched blocks contributing their base value (the value the reference recomputes).
#pragma once
#include "centralized.cuh"

namespace mas_b200 {

constexpr int kMixedMaxBlockDim = 8;  // largest state / control dimension of a registered model (checked per model below)

struct MixedBlock {
  int model_id, state_offset, control_offset;
  int nx, nu;
  double params[kMaxParams];
};

#define MAS_MIXED_SWITCH(id, ...)                          \
  switch (id) {                                            \
    case StLane::ID: { using M = StLane; __VA_ARGS__; } break;       \
    case StCirc::ID: { using M = StCirc; __VA_ARGS__; } break;       \
    case Lqr4::ID: { using M = Lqr4; __VA_ARGS__; } break;           \
    case Pendulum::ID: { using M = Pendulum; __VA_ARGS__; } break;   \
    case Rocket::ID: { using M = Rocket; __VA_ARGS__; } break;       \
    case StLaneCon::ID: { using M = StLaneCon; __VA_ARGS__; } break; \
    default: break;                                        \
  }

static_assert(StLane::NX <= kMixedMaxBlockDim && StCirc::NX <= kMixedMaxBlockDim && Lqr4::NX <= kMixedMaxBlockDim && Pendulum::NX <= kMixedMaxBlockDim &&
                  Rocket::NX <= kMixedMaxBlockDim && Lqr4::NU <= kMixedMaxBlockDim,
              "kMixedMaxBlockDim is too small for a registered model");

inline bool mixed_model_dims(int model_id, int* nx, int* nu) {
  bool found = false;
  MAS_MIXED_SWITCH(model_id, *nx = M::NX; *nu = M::NU; found = true)
  return found;
}

// The model calls are real calls on the device (six models behind every switch, dozens of call sites in the finite
// differences: inlined, the kernel would be megabytes of code).
#if defined(__CUDACC__)
#define MAS_HD_NI static __host__ __device__ __noinline__
#else
#define MAS_HD_NI inline
#endif

// the three functions an agent contributes to the stacked OCP, on its own block (x, u are the block's coordinates)
MAS_HD_NI void mixed_dynamics(int model_id, const double* prm, const double* x, const double* u, double* d) {
  MAS_MIXED_SWITCH(model_id, M::dynamics(x, u, prm, d))
}
MAS_HD_NI double mixed_stage(int model_id, const double* prm, const double* x, const double* u, int t) {
  double c = 0.0;
  MAS_MIXED_SWITCH(model_id, c = M::stage(x, u, t, prm))
  return c;
}
MAS_HD_NI double mixed_terminal(int model_id, const double* prm, const double* x) {
  double c = 0.0;
  MAS_MIXED_SWITCH(model_id, c = M::terminal(x, prm))
  return c;
}
MAS_HD_NI void mixed_rk4(int model_id, const double* prm, const double* x, const double* u, double dt, double* xn) {
  MAS_MIXED_SWITCH(model_id, rk4_step<M>(x, u, prm, dt, xn))
}

struct MixedStacked {
  int n_blocks, ns, ms, T;
  double dt;
  int has_bounds;
  double tolerance;
  int max_iterations;
  double max_ms;
  const MixedBlock* blocks;  // [n_blocks], block order; `params` are the defaults, `prm` overrides them per scenario
  const int* block_of_x;     // [ns] block index of every stacked state coordinate
  const int* block_of_u;     // [ms]
  const double* lo;          // [ms] stacked bounds (has_bounds only)
  const double* hi;
  // per scenario (already offset)
  const double* x0;   // [ns]
  const double* prm;  // [n_blocks][kMaxParams]
  double* X;          // [(T+1)*ns] best states, column t at t*ns
  double* U;          // [T*ms]
  double* work;       // MixedWork layout
  double* out_cost;   // [1 + n_blocks]: stacked best_cost, then every block's own objective (centralized.hpp:32)
  int* out_int;       // iterations, status, reg_retries, alpha_trials
};

struct MixedWork {
  size_t Xt, Ut, K, kff, Vx, Vxx, Vnew, A, B, lx, lu, lxx, luu, lux, Qx, Qu, Qxx, Qux, Quu, Qreg, L, inv, AtV, BtV, KtQuu, f0, cb, stg, dx, scal, total;
  MAS_HD MixedWork(int ns, int ms, int T, int nb) {
    size_t o = 0;
    auto take = [&](size_t n) {
      const size_t r = o;
      o += n;
      return r;
    };
    const size_t n = ns, m = ms;
    Xt = take((T + 1) * n);
    Ut = take(T * m);
    K = take(static_cast<size_t>(T) * m * n);
    kff = take(T * m);
    Vx = take(n);
    Vxx = take(n * n);
    Vnew = take(n * n);
    A = take(n * n);
    B = take(n * m);
    lx = take(n);
    lu = take(m);
    lxx = take(n * n);
    luu = take(m * m);
    lux = take(m * n);
    Qx = take(n);
    Qu = take(m);
    Qxx = take(n * n);
    Qux = take(m * n);
    Quu = take(m * m);
    Qreg = take(m * m);
    L = take(m * m);
    inv = take(m * m);
    AtV = take(n * n);
    BtV = take(m * n);
    KtQuu = take(n * m);
    f0 = take(n);
    cb = take(nb);
    stg = take(T + 1);
    dx = take(n + m);
    scal = take(8);
    total = o;
  }
};
enum MixedScalar { MS_COST = 0, MS_MERIT, MS_TRIAL, MS_FLAG, MS_C0 };

MAS_HD double mixed_safe(double v) { return isfinite(v) ? v : 0.0; }  // finite_differences.hpp:95-107

// up to two perturbed stacked coordinates: states (is_u = 0) or controls (is_u = 1)
struct MixedPert {
  int n;
  int is_u[2], idx[2];
  double delta[2];
};
MAS_HD MixedPert mixed_pert1(int is_u, int idx, double d) { return MixedPert{1, {is_u, 0}, {idx, 0}, {d, 0.0}}; }
MAS_HD MixedPert mixed_pert2(int is_u0, int i0, double d0, int is_u1, int i1, double d1) { return MixedPert{2, {is_u0, is_u1}, {i0, i1}, {d0, d1}}; }

// block a's coordinates of (x, u) with the perturbations that fall into it applied; returns whether any did
MAS_HD bool mixed_block_point(const MixedStacked& P, int a, const double* x, const double* u, const MixedPert& pt, double* xb, double* ub) {
  const MixedBlock& b = P.blocks[a];
  for (int i = 0; i < b.nx; ++i) xb[i] = x[b.state_offset + i];
  if (u)
    for (int i = 0; i < b.nu; ++i) ub[i] = u[b.control_offset + i];
  bool touched = false;
  for (int k = 0; k < pt.n; ++k) {
    if (pt.is_u[k]) {
      const int r = pt.idx[k] - b.control_offset;
      if (r >= 0 && r < b.nu) {
        ub[r] = ub[r] + pt.delta[k];
        touched = true;
      }
    } else {
      const int r = pt.idx[k] - b.state_offset;
      if (r >= 0 && r < b.nx) {
        xb[r] = xb[r] + pt.delta[k];
        touched = true;
      }
    }
  }
  return touched;
}

// stacked stage cost (multi_agent_problem.hpp:104-113) at a perturbed point: cost = 0.0; cost += block terms in order.
// cb = the blocks' terms at the unperturbed point.
MAS_HD_NI double mixed_stage_at(const MixedStacked& P, const double* cb, const double* x, const double* u, int t, const MixedPert& pt) {
  double s = 0.0;
  for (int a = 0; a < P.n_blocks; ++a) {
    double xb[kMixedMaxBlockDim], ub[kMixedMaxBlockDim];
    const bool touched = mixed_block_point(P, a, x, u, pt, xb, ub);
    s += touched ? mixed_stage(P.blocks[a].model_id, P.prm + a * kMaxParams, xb, ub, t) : cb[a];
  }
  return s;
}
MAS_HD_NI double mixed_terminal_at(const MixedStacked& P, const double* cb, const double* x, const MixedPert& pt) {
  double s = 0.0;
  for (int a = 0; a < P.n_blocks; ++a) {
    double xb[kMixedMaxBlockDim], ub[kMixedMaxBlockDim];
    const bool touched = mixed_block_point(P, a, x, nullptr, pt, xb, ub);
    s += touched ? mixed_terminal(P.blocks[a].model_id, P.prm + a * kMaxParams, xb) : cb[a];
  }
  return s;
}

// dense column-major products with the reference's summation order: k ascending, the first product starts the sum
MAS_HD double mixed_dot_tn(const double* a, int lda, int i, const double* b, int ldb, int j, int kd) {  // sum_k a(k,i) b(k,j)
  double s = a[static_cast<size_t>(i) * lda] * b[static_cast<size_t>(j) * ldb];
  for (int k = 1; k < kd; ++k) s = s + a[k + static_cast<size_t>(i) * lda] * b[k + static_cast<size_t>(j) * ldb];
  return s;
}
MAS_HD double mixed_dot_nn(const double* a, int lda, int i, const double* b, int ldb, int j, int kd) {  // sum_k a(i,k) b(k,j)
  double s = a[i] * b[static_cast<size_t>(j) * ldb];
  for (int k = 1; k < kd; ++k) s = s + a[i + static_cast<size_t>(k) * lda] * b[k + static_cast<size_t>(j) * ldb];
  return s;
}

// ---- rollout: X(:,0) = x0, then per step the trial control (alpha < 0: the given U as it is) and one RK4 step per block;
// stage values per step in stg[], the trajectory cost (ocp.hpp:14-28: sum over t from 0.0, then the terminal) in MS_TRIAL.
MAS_HD void mixed_rollout(const MixedStacked& P, const MixedWork& W, double alpha, double* Xo, double* Uo, int tid, int nthr) {
  const int ns = P.ns, ms = P.ms, T = P.T, nb = P.n_blocks;
  double* w = P.work;
  double* stg = w + W.stg;
  double* dx = w + W.dx;
  double* cbt = w + W.cb;
  for (int i = tid; i < ns; i += nthr) Xo[i] = P.x0[i];
  MAS_CTA_SYNC();
  for (int t = 0; t < T; ++t) {
    double* xt = Xo + static_cast<size_t>(t) * ns;
    double* ut = Uo + static_cast<size_t>(t) * ms;
    if (alpha >= 0.0) {  // ilqr.hpp:206-214
      const double* xn = P.X + static_cast<size_t>(t) * ns;
      const double* un = P.U + static_cast<size_t>(t) * ms;
      const double* Kt = w + W.K + static_cast<size_t>(t) * ms * ns;
      const double* kt = w + W.kff + static_cast<size_t>(t) * ms;
      for (int i = tid; i < ns; i += nthr) dx[i] = xt[i] - xn[i];
      MAS_CTA_SYNC();
      for (int i = tid; i < ms; i += nthr) {
        const double kdx = mixed_dot_nn(Kt, ms, i, dx, ns, 0, ns);
        double ui = (un[i] + alpha * kt[i]) + kdx;
        if (P.has_bounds) {  // clamp_controls: cwiseMin(upper) then cwiseMax(lower)
          ui = (P.hi[i] < ui) ? P.hi[i] : ui;
          ui = (P.lo[i] > ui) ? P.lo[i] : ui;
        }
        ut[i] = ui;
      }
      MAS_CTA_SYNC();
    }
    for (int a = tid; a < nb; a += nthr) {
      const MixedBlock& b = P.blocks[a];
      const double* pa = P.prm + a * kMaxParams;
      cbt[a] = mixed_stage(b.model_id, pa, xt + b.state_offset, ut + b.control_offset, t);
      mixed_rk4(b.model_id, pa, xt + b.state_offset, ut + b.control_offset, P.dt, xt + ns + b.state_offset);
    }
    MAS_CTA_SYNC();
    if (tid == 0) {
      double s = 0.0;
      for (int a = 0; a < nb; ++a) s += cbt[a];
      stg[t] = s;
    }
    MAS_CTA_SYNC();
  }
  const double* xT = Xo + static_cast<size_t>(T) * ns;
  for (int a = tid; a < nb; a += nthr) cbt[a] = mixed_terminal(P.blocks[a].model_id, P.prm + a * kMaxParams, xT + P.blocks[a].state_offset);
  MAS_CTA_SYNC();
  if (tid == 0) {
    double term = 0.0;
    for (int a = 0; a < nb; ++a) term += cbt[a];
    double c = 0.0;
    for (int t = 0; t < T; ++t) c += stg[t];
    c += term;
    w[W.scal + MS_TRIAL] = c;
  }
  MAS_CTA_SYNC();
}

// `m = 0.5 * (m + m.transpose())` evaluated in place, columns outer (ilqr.hpp:102,192): the strict lower triangle averages
// two old entries, the strict upper triangle reads the lower entry that has already been updated.  src -> dst.
MAS_HD void mixed_symmetrize(const double* src, double* dst, int n, int tid, int nthr) {
  for (int e = tid; e < n * n; e += nthr) {
    const int i = e % n, j = e / n;
    const double a = src[i + static_cast<size_t>(j) * n], b = src[j + static_cast<size_t>(i) * n];
    const double lower = 0.5 * (b + a);  // the updated entry (j, i) when j > i
    dst[e] = (i < j) ? 0.5 * (a + lower) : 0.5 * (a + b);
  }
  MAS_CTA_SYNC();
}

// ---- backward pass (ilqr.hpp:92-193) on the stacked OCP, every derivative by the finite-difference defaults
// (finite_differences.hpp:53-287).  Returns nothing; K, kff in the workspace, retries added to out_int[2].
MAS_HD void mixed_backward(const MixedStacked& P, const MixedWork& W, int tid, int nthr) {
  const int ns = P.ns, ms = P.ms, T = P.T, nb = P.n_blocks;
  const double e6 = 1e-6, e5 = 1e-5;
  double* w = P.work;
  double *Vx = w + W.Vx, *Vxx = w + W.Vxx, *Vnew = w + W.Vnew, *A = w + W.A, *B = w + W.B, *lx = w + W.lx, *lu = w + W.lu, *lxx = w + W.lxx,
         *luu = w + W.luu, *lux = w + W.lux, *Qx = w + W.Qx, *Qu = w + W.Qu, *Qxx = w + W.Qxx, *Qux = w + W.Qux, *Quu = w + W.Quu, *Qreg = w + W.Qreg,
         *L = w + W.L, *inv = w + W.inv, *AtV = w + W.AtV, *BtV = w + W.BtV, *KtQuu = w + W.KtQuu, *f0 = w + W.f0, *cb = w + W.cb, *scal = w + W.scal;
  // terminal value (:92-102): FD gradient (eps 1e-6, no safe_eval) and Hessian (eps 1e-5, safe_eval) of the stacked terminal cost
  const double* xT = P.X + static_cast<size_t>(T) * ns;
  for (int a = tid; a < nb; a += nthr) cb[a] = mixed_terminal(P.blocks[a].model_id, P.prm + a * kMaxParams, xT + P.blocks[a].state_offset);
  MAS_CTA_SYNC();
  if (tid == 0) {
    double s = 0.0;
    for (int a = 0; a < nb; ++a) s += cb[a];
    scal[MS_C0] = s;
  }
  MAS_CTA_SYNC();
  for (int i = tid; i < ns; i += nthr) {
    const double fp = mixed_terminal_at(P, cb, xT, mixed_pert1(0, i, e6)), fm = mixed_terminal_at(P, cb, xT, mixed_pert1(0, i, -e6));
    Vx[i] = MAS_DIV_CONST(fp - fm, 2 * e6);
  }
  for (int e = tid; e < ns * ns; e += nthr) {
    const int i = e % ns, j = e / ns;
    double h;
    if (i == j) {
      const double fp = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert1(0, i, e5))), f00 = mixed_safe(scal[MS_C0]),
                   fm = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert1(0, i, -e5)));
      h = MAS_DIV_CONST(fp - 2 * f00 + fm, e5 * e5);
    } else {
      const double fpp = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert2(0, i, e5, 0, j, e5))),
                   fpm = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert2(0, i, e5, 0, j, -e5))),
                   fmp = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert2(0, i, -e5, 0, j, e5))),
                   fmm = mixed_safe(mixed_terminal_at(P, cb, xT, mixed_pert2(0, i, -e5, 0, j, -e5)));
      h = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * e5 * e5);
    }
    Vnew[e] = h;
  }
  MAS_CTA_SYNC();
  mixed_symmetrize(Vnew, Vxx, ns, tid, nthr);

  for (int t = T - 1; t >= 0; --t) {
    const double* x = P.X + static_cast<size_t>(t) * ns;
    const double* u = P.U + static_cast<size_t>(t) * ms;
    // base values: every block's stage term and dynamics at (x_t, u_t)
    for (int a = tid; a < nb; a += nthr) {
      const MixedBlock& b = P.blocks[a];
      const double* pa = P.prm + a * kMaxParams;
      cb[a] = mixed_stage(b.model_id, pa, x + b.state_offset, u + b.control_offset, t);
      mixed_dynamics(b.model_id, pa, x + b.state_offset, u + b.control_offset, f0 + b.state_offset);
    }
    MAS_CTA_SYNC();
    if (tid == 0) {
      double s = 0.0;
      for (int a = 0; a < nb; ++a) s += cb[a];
      scal[MS_C0] = s;
    }
    MAS_CTA_SYNC();
    // A, B (:53-92): column i perturbs one stacked coordinate; rows outside its block see equal values on both sides
    for (int c = tid; c < ns + ms; c += nthr) {
      const bool is_u = c >= ns;
      const int i = is_u ? c - ns : c;
      const int a = is_u ? P.block_of_u[i] : P.block_of_x[i];
      const MixedBlock& b = P.blocks[a];
      const double* pa = P.prm + a * kMaxParams;
      double xb[kMixedMaxBlockDim], ub[kMixedMaxBlockDim], fp[kMixedMaxBlockDim], fm[kMixedMaxBlockDim];
      mixed_block_point(P, a, x, u, mixed_pert1(is_u, i, e6), xb, ub);
      mixed_dynamics(b.model_id, pa, xb, ub, fp);
      mixed_block_point(P, a, x, u, mixed_pert1(is_u, i, -e6), xb, ub);
      mixed_dynamics(b.model_id, pa, xb, ub, fm);
      double* col = (is_u ? B : A) + static_cast<size_t>(i) * ns;
      for (int r = 0; r < ns; ++r) {
        const int rb = r - b.state_offset;
        const double diff = (rb >= 0 && rb < b.nx) ? fp[rb] - fm[rb] : f0[r] - f0[r];
        col[r] = MAS_DIV_CONST(diff, 2 * e6);
      }
    }
    // gradients (:110-136, no safe_eval)
    for (int c = tid; c < ns + ms; c += nthr) {
      const bool is_u = c >= ns;
      const int i = is_u ? c - ns : c;
      const double fp = mixed_stage_at(P, cb, x, u, t, mixed_pert1(is_u, i, e6)), fm = mixed_stage_at(P, cb, x, u, t, mixed_pert1(is_u, i, -e6));
      (is_u ? lu : lx)[i] = MAS_DIV_CONST(fp - fm, 2 * e6);
    }
    // Hessians (:138-210, eps 1e-5, safe_eval) and the cross term (:263-287, eps 1e-6, f_pm = x - eps, u + eps)
    for (int e = tid; e < ns * ns + ms * ms + ms * ns; e += nthr) {
      double h;
      if (e < ns * ns + ms * ms) {
        const bool is_u = e >= ns * ns;
        const int n = is_u ? ms : ns, f = is_u ? e - ns * ns : e, i = f % n, j = f / n;
        if (i == j) {
          const double fp = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert1(is_u, i, e5))), f00 = mixed_safe(scal[MS_C0]),
                       fm = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert1(is_u, i, -e5)));
          h = MAS_DIV_CONST(fp - 2 * f00 + fm, e5 * e5);
        } else {
          const double fpp = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(is_u, i, e5, is_u, j, e5))),
                       fpm = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(is_u, i, e5, is_u, j, -e5))),
                       fmp = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(is_u, i, -e5, is_u, j, e5))),
                       fmm = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(is_u, i, -e5, is_u, j, -e5)));
          h = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * e5 * e5);
        }
        (is_u ? luu : lxx)[f] = h;
      } else {
        const int f = e - ns * ns - ms * ms, i = f % ms, j = f / ms;  // l_ux(i, j): control i, state j
        const double fpp = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(0, j, e6, 1, i, e6))),
                     fpm = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(0, j, -e6, 1, i, e6))),
                     fmp = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(0, j, e6, 1, i, -e6))),
                     fmm = mixed_safe(mixed_stage_at(P, cb, x, u, t, mixed_pert2(0, j, -e6, 1, i, -e6)));
        lux[f] = MAS_DIV_CONST(fpp - fpm - fmp + fmm, 4 * e6 * e6);
      }
    }
    MAS_CTA_SYNC();
    // Q assembly (:115-119): (A^T V_xx) and (B^T V_xx) first, then multiplied on the right
    for (int e = tid; e < ns + ms + ns * ns + ms * ns; e += nthr) {
      if (e < ns) {
        Qx[e] = lx[e] + mixed_dot_tn(A, ns, e, Vx, ns, 0, ns);
      } else if (e < ns + ms) {
        const int i = e - ns;
        Qu[i] = lu[i] + mixed_dot_tn(B, ns, i, Vx, ns, 0, ns);
      } else if (e < ns + ms + ns * ns) {
        const int f = e - ns - ms;
        AtV[f] = mixed_dot_tn(A, ns, f % ns, Vxx, ns, f / ns, ns);
      } else {
        const int f = e - ns - ms - ns * ns;
        BtV[f] = mixed_dot_tn(B, ns, f % ms, Vxx, ns, f / ms, ns);
      }
    }
    MAS_CTA_SYNC();
    for (int e = tid; e < ns * ns + ms * ns + ms * ms; e += nthr) {
      if (e < ns * ns) {
        Qxx[e] = lxx[e] + mixed_dot_nn(AtV, ns, e % ns, A, ns, e / ns, ns);
      } else if (e < ns * ns + ms * ns) {
        const int f = e - ns * ns;
        Qux[f] = lux[f] + mixed_dot_nn(BtV, ms, f % ms, A, ns, f / ms, ns);
      } else {
        const int f = e - ns * ns - ms * ns;
        const double q = luu[f] + mixed_dot_nn(BtV, ms, f % ms, B, ns, f / ms, ns);
        Quu[f] = q;
        Qreg[f] = q;
      }
    }
    MAS_CTA_SYNC();
    // Q_uu + reg I until the unblocked lower LLT succeeds (:172-183); one thread: ms is a handful of controls per agent
    if (tid == 0) {
      double reg = 1e-6;
      int retries = 0;
      for (;;) {
        bool ok = true;
        for (int i = 0; i < ms * ms; ++i) L[i] = Qreg[i];
        for (int k = 0; k < ms && ok; ++k) {
          double xk = L[k + static_cast<size_t>(k) * ms];
          if (k > 0) {
            double sq = 0.0;
            for (int j = 0; j < k; ++j) sq += L[k + static_cast<size_t>(j) * ms] * L[k + static_cast<size_t>(j) * ms];
            xk -= sq;
          }
          if (xk <= 0.0) {
            ok = false;
          } else {
            xk = sqrt(xk);
            L[k + static_cast<size_t>(k) * ms] = xk;
            for (int i = k + 1; i < ms; ++i) {
              double s = L[i + static_cast<size_t>(k) * ms];
              if (k > 0) {
                double acc = L[i] * L[k];
                for (int j = 1; j < k; ++j) acc = acc + L[i + static_cast<size_t>(j) * ms] * L[k + static_cast<size_t>(j) * ms];
                s -= acc;
              }
              L[i + static_cast<size_t>(k) * ms] = pm::div_(s, xk);
            }
          }
        }
        if (ok) break;
        for (int i = 0; i < ms; ++i) Qreg[i + static_cast<size_t>(i) * ms] += reg;
        reg *= 10.0;
        ++retries;
        if (!(reg < 1e300)) break;  // as in riccati_step: the reference would loop forever
      }
      P.out_int[2] += retries;
    }
    MAS_CTA_SYNC();
    // Q_uu_inv = llt.solve(I): forward and back substitution, one column per thread
    for (int c = tid; c < ms; c += nthr) {
      double* xc = inv + static_cast<size_t>(c) * ms;
      for (int i = 0; i < ms; ++i) xc[i] = (i == c) ? 1.0 : 0.0;
      for (int i = 0; i < ms; ++i) {
        double s = xc[i];
        for (int j = 0; j < i; ++j) s -= L[i + static_cast<size_t>(j) * ms] * xc[j];
        xc[i] = pm::div_(s, L[i + static_cast<size_t>(i) * ms]);
      }
      for (int i = ms - 1; i >= 0; --i) {
        double s = xc[i];
        for (int j = i + 1; j < ms; ++j) s -= L[j + static_cast<size_t>(i) * ms] * xc[j];
        xc[i] = pm::div_(s, L[i + static_cast<size_t>(i) * ms]);
      }
    }
    MAS_CTA_SYNC();
    for (int e = tid; e < ms * ms; e += nthr) L[e] = -inv[e];  // L now holds -Q_uu_inv (:185-186)
    MAS_CTA_SYNC();
    double* Kt = w + W.K + static_cast<size_t>(t) * ms * ns;
    double* kt = w + W.kff + static_cast<size_t>(t) * ms;
    for (int e = tid; e < ms + ms * ns; e += nthr) {
      if (e < ms) kt[e] = mixed_dot_nn(L, ms, e, Qu, ms, 0, ms);
      else Kt[e - ms] = mixed_dot_nn(L, ms, (e - ms) % ms, Qux, ms, (e - ms) / ms, ms);
    }
    MAS_CTA_SYNC();
    // value update with the unregularised Q_uu (:188-192)
    for (int e = tid; e < ns * ms; e += nthr) KtQuu[e] = mixed_dot_tn(Kt, ms, e % ns, Quu, ms, e / ns, ms);
    MAS_CTA_SYNC();
    for (int e = tid; e < ns + ns * ns; e += nthr) {
      if (e < ns) {
        const double t1 = mixed_dot_tn(Kt, ms, e, Qu, ms, 0, ms), t2 = mixed_dot_tn(Qux, ms, e, kt, ms, 0, ms), t3 = mixed_dot_nn(KtQuu, ns, e, kt, ms, 0, ms);
        Vx[e] = ((Qx[e] + t1) + t2) + t3;
      } else {
        const int f = e - ns, i = f % ns, j = f / ns;
        const double m1 = mixed_dot_tn(Kt, ms, i, Qux, ms, j, ms), m2 = mixed_dot_tn(Qux, ms, i, Kt, ms, j, ms), m3 = mixed_dot_nn(KtQuu, ns, i, Kt, ms, j, ms);
        Vnew[f] = ((Qxx[f] + m1) + m2) + m3;
      }
    }
    MAS_CTA_SYNC();
    mixed_symmetrize(Vnew, Vxx, ns, tid, nthr);
  }
}

// iLQR::solve on the stacked OCP (ilqr.hpp:59-273; no constraints: build_global_ocp does not stack them), then every
// block's own objective on its rows of the result (centralized.hpp:27-36).
MAS_HD void mixed_stacked_solve(const MixedStacked& P, int tid, int nthr) {
  const MixedWork W(P.ns, P.ms, P.T, P.n_blocks);
  const int ns = P.ns, ms = P.ms, T = P.T;
  double* w = P.work;
  double* scal = w + W.scal;
  double *Xt = w + W.Xt, *Ut = w + W.Ut;
  if (tid == 0) {
    P.out_int[0] = 0;
    P.out_int[1] = STATUS_MAX_ITER;
    P.out_int[2] = 0;
    P.out_int[3] = 0;
  }
  for (int i = tid; i < T * ms; i += nthr) P.U[i] = 0.0;  // the stacked OCP starts from zero controls (ocp.hpp:104-108)
  MAS_CTA_SYNC();
  mixed_rollout(P, W, -1.0, P.X, P.U, tid, nthr);
  if (tid == 0) {
    scal[MS_COST] = scal[MS_TRIAL];
    scal[MS_MERIT] = scal[MS_TRIAL];
  }
  MAS_CTA_SYNC();
  const bool timed = P.max_ms < 1.7976931348623157e308;
  const unsigned long long start_ns = timed ? stacked_now_ns() : 0ull;
  for (int iter = 0; iter < P.max_iterations; ++iter) {
    if (timed) {  // whole milliseconds, checked only here (ilqr.hpp:84-90)
      if (tid == 0) scal[MS_FLAG] = static_cast<double>((stacked_now_ns() - start_ns) / 1000000ull) > P.max_ms ? 1.0 : 0.0;
      MAS_CTA_SYNC();
      const bool out_of_time = scal[MS_FLAG] != 0.0;
      MAS_CTA_SYNC();
      if (out_of_time) {
        if (tid == 0) P.out_int[1] = STATUS_TIME_LIMIT;
        break;
      }
    }
    if (tid == 0) P.out_int[0] = iter + 1;
    mixed_backward(P, W, tid, nthr);
    const double current_merit = scal[MS_MERIT];
    int accepted = -1;
    double best_merit = current_merit, alpha = 1.0;
    for (int j = 0; j < kNumAlphas; ++j) {  // alpha = 1, 1/2, ... >= 1e-3; the first improvement wins (:195-228)
      mixed_rollout(P, W, alpha, Xt, Ut, tid, nthr);
      const double trial = scal[MS_TRIAL];
      MAS_CTA_SYNC();
      if (tid == 0) P.out_int[3] += 1;
      if (trial < best_merit) {
        best_merit = trial;
        accepted = j;
        break;
      }
      alpha *= 0.5;
    }
    if (accepted >= 0) {
      for (int i = tid; i < (T + 1) * ns; i += nthr) P.X[i] = Xt[i];
      for (int i = tid; i < T * ms; i += nthr) P.U[i] = Ut[i];
      if (tid == 0) {
        scal[MS_COST] = best_merit;  // the objective of the accepted trajectory is the same sum (:233)
        scal[MS_MERIT] = best_merit;
      }
    }
    MAS_CTA_SYNC();
    if (current_merit - best_merit < P.tolerance) {  // :269-271 (NaN improvement: keeps iterating, like the reference)
      if (tid == 0) P.out_int[1] = STATUS_CONVERGED;
      break;
    }
  }
  MAS_CTA_SYNC();
  for (int a = tid; a < P.n_blocks; a += nthr) {
    const MixedBlock& b = P.blocks[a];
    const double* pa = P.prm + a * kMaxParams;
    double c = 0.0;
    for (int t = 0; t < T; ++t) c += mixed_stage(b.model_id, pa, P.X + static_cast<size_t>(t) * ns + b.state_offset, P.U + static_cast<size_t>(t) * ms + b.control_offset, t);
    c += mixed_terminal(b.model_id, pa, P.X + static_cast<size_t>(T) * ns + b.state_offset);
    P.out_cost[1 + a] = c;
  }
  if (tid == 0) P.out_cost[0] = scal[MS_COST];
}

}  // namespace mas_b200
