// host_emulation.cpp -- TEST HARNESS, not a product path.
//
// Compiles the per-thread bodies of the CUDA kernels (multi_agent_solver_b200/csrc/ilqr_core.cuh,
// models.cuh -- all __host__ __device__) with g++ and drives them with the same schedule the engine
// uses on the GPU (prologue, then per iteration: backward for each active problem, L-lane line
// search with the group reduction done by a loop instead of warp shuffles, commit, stop test,
// active-list compaction).  This lets the "-m 'not gpu'" tests check the device source against the
// oracle bit for bit on a machine without a GPU.  Nothing in the product links or loads this file.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "ilqr_core.cuh"

using namespace mas_b200;

namespace {

// Constrained models: solver-object state that outlives a solve (BatchBase::prepare_constraint_state) and the
// repeat count for "the same solver solving again"; hist_* receive [repeats][batch] when non-null.
struct ALOptions {
  double penalty = 10.0, penalty_increase = 5.0, constraint_tolerance = 1e-4, activation_tolerance = 1e-6;
  int repeats = 1;
  bool trial_store = true;
  bool sweep_lanes = false;  // time-parallel mode: the Riccati sweep column-parallel over the lanes of a problem (RiccatiLanes)
  int sweep_wide = 0;        // time-parallel mode: one lane per matrix entry (RiccatiWide); 1 / 2 = lanes run a phase in ascending / mixed order
  int backward_lanes = 0;  // > 0: lane-parallel backward pass with that many lanes per problem; < 0: time-parallel with -n threads per point
  double* hist_cost = nullptr;
  int* hist_iters = nullptr;
};
ALOptions g_al;

// backward_lanes (ilqr_core.cuh) with the lanes of a group run one after the other inside every phase: the FD tasks of
// a step by "lane" 0..LB-1, then lane 0's gather + riccati_step.  Same task functions, same block layout.
template <class M, int MASK_CT>
int emulate_backward_lanes(const BatchView<M::NX, M::NU>& v, int p, int LB) {
  constexpr int NX = M::NX, NU = M::NU;
  using D = DerivBlock<M>;
  const unsigned mask = (MASK_CT >= 0) ? static_cast<unsigned>(MASK_CT) : v.deriv_mask;
  double prm[M::NP > 0 ? M::NP : 1];
  load_params<M>(v, p, prm);
  const int T = v.T;
  int retries = 0;
  const double al_rho = HasConstraints<M>::value ? v.penalty[p] : 0.0;
  std::vector<double> blk(D::size > D::n_terminal_tasks ? D::size : D::n_terminal_tasks, 0.0);
  double x[NX], u[NU], v_x[NX], v_xx[NX * NX];
  for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(T, i, v.ld, p)];
  for (int lane = 0; lane < LB; ++lane)
    for (int task = lane; task < D::n_terminal_tasks; task += LB) fd_terminal_task<M>(mask, task, x, prm, blk.data());
  if (mask & D_VX) M::v_x(x, prm, v_x);
  else
    for (int i = 0; i < NX; ++i) v_x[i] = blk[i];
  if (mask & D_VXX) M::v_xx(x, prm, v_xx);
  else
    for (int i = 0; i < NX * NX; ++i) v_xx[i] = blk[NX + i];
  symmetrize_aliased<NX>(v_xx);
  for (int t = T - 1; t >= 0; --t) {
    for (int i = 0; i < NX; ++i) x[i] = v.X[soa_index<NX>(t, i, v.ld, p)];
    for (int i = 0; i < NU; ++i) u[i] = v.U[soa_index<NU>(t, i, v.ld, p)];
    for (int lane = 0; lane < LB; ++lane)
      for (int task = lane; task < D::n_tasks; task += LB) fd_stage_task<M>(mask, task, x, u, t, prm, blk.data());
    double A[NX * NX], B[NX * NU], l_x[NX], l_u[NU], l_xx[NX * NX], l_uu[NU * NU], l_ux[NU * NX];
    gather_stage_derivatives<M>(mask, blk.data(), x, u, t, prm, A, B, l_x, l_u, l_xx, l_uu, l_ux);
    retries += riccati_step<M, MASK_CT>(v, p, t, x, u, prm, al_rho, A, B, l_x, l_u, l_xx, l_uu, l_ux, v_x, v_xx);
  }
  return retries;
}

// Time-parallel backward pass (linearize_kernel + riccati_sweep_kernel): the derivative blocks of all T + 1 points are
// produced first, each by G emulated threads, then the sweep consumes them.
template <class M, int MASK_CT>
int emulate_time_parallel(const BatchView<M::NX, M::NU>& v, int p, int G) {
  using D = DerivBlock<M>;
  const unsigned mask = (MASK_CT >= 0) ? static_cast<unsigned>(MASK_CT) : v.deriv_mask;
  std::vector<double> store(static_cast<size_t>(v.T + 1) * D::size, 0.0);
  for (int t = v.T; t >= 0; --t)
    for (int g = G - 1; g >= 0; --g)
      linearize_point<M>(v, p, t, mask, g, G, [&](int off, double val) { store[static_cast<size_t>(t) * D::size + off] = val; });
  if constexpr (RiccatiWide<M, MASK_CT>::kSupported) {
    if (g_al.sweep_wide) {
      // riccati_sweep_wide_kernel: the LW lanes of the problem run every phase one after the other (a barrier separates the
      // phases on the device), in an order that changes from phase to phase so that a dependence inside a phase would show
      using RW = RiccatiWide<M, MASK_CT>;
      constexpr int LW = RW::LW;
      RW lane[LW];
      double xch[RW::XCH];
      for (int k = 0; k < RW::XCH; ++k) xch[k] = std::numeric_limits<double>::quiet_NaN();
      int phase = 0;
      auto each = [&](auto fn) {
        ++phase;
        for (int k = 0; k < LW; ++k) {
          const int e = g_al.sweep_wide == 1 ? k : ((phase & 1) ? LW - 1 - k : (k * 5 + phase) % LW);
          fn(e);
        }
      };
      const double* tb = &store[static_cast<size_t>(v.T) * D::size];
      each([&](int e) { RW::init_terminal(tb, e, xch); });
      each([&](int e) { RW::phase6(e, xch); });
      for (int t = v.T - 1; t >= 0; --t) {
        const double* blk = &store[static_cast<size_t>(t) * D::size];
        each([&](int e) { lane[e].phase1(blk, e, xch); });
        each([&](int e) { lane[e].phase2(blk, e, xch); });
        each([&](int e) { lane[e].phase3(v, p, t, e, xch); });
        each([&](int e) { lane[e].phase4(e, xch); });
        each([&](int e) { lane[e].phase5(e, xch); });
        each([&](int e) { RW::phase6(e, xch); });
      }
      return lane[0].retries;
    }
  }
  if (g_al.sweep_lanes && !HasConstraints<M>::value) {
    // riccati_sweep_lanes_kernel: the lanes of the problem run every phase one after the other, exchange area in between
    using RL = RiccatiLanes<M, MASK_CT>;
    RL lane[M::NX];
    double xch[RL::XCH];
    for (int j = 0; j < M::NX; ++j) lane[j].init_terminal(&store[static_cast<size_t>(v.T) * D::size], j);
    for (int t = v.T - 1; t >= 0; --t) {
      const double* blk = &store[static_cast<size_t>(t) * D::size];
      for (int j = M::NX - 1; j >= 0; --j) lane[j].phase_a(blk, j, xch);
      for (int j = 0; j < M::NX; ++j) lane[j].phase_b(blk, j, xch);
      for (int j = M::NX - 1; j >= 0; --j) lane[j].phase_c(j, xch);
      for (int j = 0; j < M::NX; ++j) lane[j].phase_d(v, p, t, j, xch);
      for (int j = M::NX - 1; j >= 0; --j) lane[j].phase_e(j, xch);
    }
    return lane[0].retries;
  }
  return riccati_sweep_thread<M, MASK_CT>(v, p, [&](int t, double* blk) {
    for (int k = 0; k < D::size; ++k) blk[k] = store[static_cast<size_t>(t) * D::size + k];
  });
}

template <class M>
int emulate(int batch, int T, double dt, unsigned mask, int has_bounds, const double* lo, const double* hi, const double* shared_p,
            const double* per_problem_p, const double* x0, double* U, double* X, double* cost, int* iters, int* status, int* trials, int* reg,
            int max_iterations, double tolerance, int L, int C) {
  constexpr int NX = M::NX, NU = M::NU;
  const int ld = ((batch + 31) / 32) * 32;
  std::vector<double> sx0(static_cast<size_t>(NX) * ld), sX(static_cast<size_t>(NX) * (T + 1) * ld), sU(static_cast<size_t>(NU) * T * ld),
      sK(static_cast<size_t>(NU) * NX * T * ld), sk(static_cast<size_t>(NU) * T * ld), scost(ld), smerit(ld), sp(static_cast<size_t>(kMaxParams) * ld);
  std::vector<int> sit(ld), sst(ld), str(ld), srg(ld);
  for (int b = 0; b < batch; ++b) {
    for (int i = 0; i < NX; ++i) sx0[static_cast<size_t>(i) * ld + b] = x0[static_cast<size_t>(b) * NX + i];
    for (int r = 0; r < NU * T; ++r) sU[static_cast<size_t>(r) * ld + b] = U[static_cast<size_t>(b) * NU * T + r];
    if (per_problem_p)
      for (int i = 0; i < M::NP; ++i) sp[static_cast<size_t>(i) * ld + b] = per_problem_p[static_cast<size_t>(b) * M::NP + i];
  }
  BatchView<NX, NU> v{};
  v.ld = ld;
  v.T = T;
  v.set_dt(dt);
  v.deriv_mask = mask;
  v.set_bounds(has_bounds, lo, hi);
  v.per_problem_params = per_problem_p ? 1 : 0;
  for (int i = 0; i < kMaxParams; ++i) v.shared_p[i] = shared_p ? shared_p[i] : 0.0;
  v.params = sp.data();
  v.x0 = sx0.data();
  v.X = sX.data();
  v.U = sU.data();
  v.K = sK.data();
  v.kff = sk.data();
  v.cost = scost.data();
  v.merit = smerit.data();
  v.iters = sit.data();
  v.status = sst.data();
  v.trials = str.data();
  v.reg_retries = srg.data();
  v.tolerance = tolerance;
  v.max_iterations = max_iterations;
  constexpr int NEQs = M::NEQ > 0 ? M::NEQ : 1, NINEQs = M::NINEQ > 0 ? M::NINEQ : 1;
  std::vector<double> slam_eq(static_cast<size_t>(NEQs) * T * ld, 0.0), slam_ineq(static_cast<size_t>(NINEQs) * T * ld, 0.0), spen(ld, g_al.penalty);
  v.lam_eq = slam_eq.data();
  v.lam_ineq = slam_ineq.data();
  v.penalty = spen.data();
  v.penalty_increase = g_al.penalty_increase;
  v.constraint_tolerance = g_al.constraint_tolerance;
  v.activation_tolerance = g_al.activation_tolerance;
  if (HasConstraints<M>::value && T > kMaxALHorizon) return 2;
  // trial store (BatchView::trial_*): 64 slots, reused by every emulated warp / lane group in turn
  std::vector<double> strial_X(static_cast<size_t>(NX) * T * 64), strial_U(static_cast<size_t>(NU) * T * 64);
  v.trial_X = g_al.trial_store ? strial_X.data() : nullptr;
  v.trial_U = g_al.trial_store ? strial_U.data() : nullptr;
  v.trial_slots = g_al.trial_store ? 64 : 0;
  const bool store = g_al.trial_store;

 for (int rep = 0; rep < g_al.repeats; ++rep) {
  std::vector<int> list(batch), next;
  for (int p = 0; p < batch; ++p) {  // prologue_kernel
    const double c = rollout_thread<M>(v, p);
    v.cost[p] = c;
    if (HasConstraints<M>::value) {
      double prm[M::NP > 0 ? M::NP : 1];
      load_params<M>(v, p, prm);
      v.merit[p] = al_merit_of_stored<M>(v, p, prm, c);
    } else {
      v.merit[p] = c;
    }
    v.iters[p] = 0;
    v.trials[p] = 0;
    v.reg_retries[p] = 0;
    v.status[p] = STATUS_MAX_ITER;
    list[p] = p;
  }
  if (max_iterations <= 0) list.clear();
  for (int it = 0; it < max_iterations && !list.empty(); ++it) {
    for (int p : list) {  // backward_kernel
      int r;
      if (g_al.backward_lanes < 0 && mask == M::EXAMPLE_MASK) r = emulate_time_parallel<M, static_cast<int>(M::EXAMPLE_MASK)>(v, p, -g_al.backward_lanes);
      else if (g_al.backward_lanes < 0 && mask == 0u) r = emulate_time_parallel<M, 0>(v, p, -g_al.backward_lanes);
      else if (g_al.backward_lanes < 0) r = emulate_time_parallel<M, -1>(v, p, -g_al.backward_lanes);
      else if (g_al.backward_lanes > 0 && mask == 0u) r = emulate_backward_lanes<M, 0>(v, p, g_al.backward_lanes);
      else if (g_al.backward_lanes > 0) r = emulate_backward_lanes<M, -1>(v, p, g_al.backward_lanes);
      else if (mask == M::EXAMPLE_MASK) r = backward_thread<M, static_cast<int>(M::EXAMPLE_MASK)>(v, p);
      else if (mask == 0u) r = backward_thread<M, 0>(v, p);
      else r = backward_thread<M, -1>(v, p);
      v.reg_retries[p] += r;
    }
    next.clear();
    if (L == 32) {  // forward_coop_kernel: a warp owns 32 consecutive list entries, lanes emulated by loops
      for (size_t first = 0; first < list.size(); first += 32) {
        const int n_valid = static_cast<int>(std::min<size_t>(32, list.size() - first));
        int done[32], nxt[32], acc[32];
        double cur[32], accm[32], merits[32][kNumAlphas];
        CoopPlan plan;
        for (int i = 0; i < 32; ++i) {
          done[i] = i < n_valid ? 0 : 1;
          nxt[i] = 0;
          acc[i] = -1;
          accm[i] = 0.0;
          cur[i] = i < n_valid ? v.merit[list[first + i]] : 0.0;
        }
        for (int round = 0; round < kNumAlphas; ++round) {
          coop_assign(done, nxt, n_valid, &plan, C == 2 ? 2 : 1);
          bool any = false;
          for (int i = 0; i < n_valid; ++i) any = any || plan.quota[i] > 0;
          if (!any) break;
          for (int lane = 0; lane < 32; ++lane) {
            const int o = plan.owner[lane];
            if (o < 0) continue;
            const int po = list[first + o];
            double prm[M::NP > 0 ? M::NP : 1];
            load_params<M>(v, po, prm);
            const int j = plan.cand[lane];
            if (C == 2) {
              double alpha[2] = {alpha_of(j), alpha_of(j + 1 < kNumAlphas ? j + 1 : kNumAlphas - 1)}, m2[2];
              trial_rollout<M, 2>(v, po, prm, alpha, m2);
              for (int c = 0; c < 2; ++c)
                if (j + c < kNumAlphas) merits[o][j + c] = m2[c];
            } else {
              const double alpha = alpha_of(j);
              trial_rollout<M, 1>(v, po, prm, &alpha, &merits[o][j]);
            }
          }
          for (int i = 0; i < n_valid; ++i)
            if (!done[i]) done[i] = coop_owner_update(merits[i], cur[i], plan.quota[i], &nxt[i], &acc[i], &accm[i], C == 2 ? 2 : 1) ? 1 : 0;
        }
        for (int i = 0; i < n_valid; ++i) {
          const int p = list[first + i];
          double prm[M::NP > 0 ? M::NP : 1];
          load_params<M>(v, p, prm);
          if (finish_iteration<M>(v, p, prm, cur[i], acc[i] >= 0 ? acc[i] : kNumAlphas, acc[i] >= 0 ? accm[i] : cur[i])) next.push_back(p);
        }
      }
      list.swap(next);
      continue;
    }
    for (int p : list) {  // forward_kernel, lanes emulated one after the other
      double prm[M::NP > 0 ? M::NP : 1];
      load_params<M>(v, p, prm);
      const double current_merit = v.merit[p];
      int best_j = kNumAlphas, best_slot = -1;
      double best_merit = 0.0, best_obj = 0.0;
      for (int lane = 0; lane < L; ++lane) {
        int bj, bs;
        double bm, bo = 0.0;
        bs = -1;  // forward_kernel keeps trial trajectories for L >= 4 only: lane l -> slot l
        if (L == 1 && C == 2) lane_line_search<M, 1, 2>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 1) lane_line_search<M, 1, 1>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 2 && C == 2) lane_line_search<M, 2, 2>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 2) lane_line_search<M, 2, 1>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 4 && C == 2 && store) lane_line_search<M, 4, 2, true>(v, p, prm, lane, current_merit, &bj, &bm, lane, 16, &bs, &bo);
        else if (L == 4 && C == 2) lane_line_search<M, 4, 2>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 4 && store) lane_line_search<M, 4, 1, true>(v, p, prm, lane, current_merit, &bj, &bm, lane, 16, &bs, &bo);
        else if (L == 4) lane_line_search<M, 4, 1>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 8 && C == 2 && store) lane_line_search<M, 8, 2, true>(v, p, prm, lane, current_merit, &bj, &bm, lane, 16, &bs, &bo);
        else if (L == 8 && C == 2) lane_line_search<M, 8, 2>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (L == 8 && store) lane_line_search<M, 8, 1, true>(v, p, prm, lane, current_merit, &bj, &bm, lane, 16, &bs, &bo);
        else if (L == 8) lane_line_search<M, 8, 1>(v, p, prm, lane, current_merit, &bj, &bm);
        else if (store) lane_line_search<M, 16, 1, true>(v, p, prm, lane, current_merit, &bj, &bm, lane, 16, &bs, &bo);
        else lane_line_search<M, 16, 1>(v, p, prm, lane, current_merit, &bj, &bm);
        if (bj < best_j) {
          best_j = bj;
          best_merit = bm;
          best_slot = bs;
          best_obj = bo;
        }
      }
      if (best_j == kNumAlphas) best_merit = current_merit;
      if (finish_iteration<M>(v, p, prm, current_merit, best_j, best_merit, best_slot, best_obj)) next.push_back(p);
    }
    list.swap(next);
  }
  for (int b = 0; b < batch; ++b) {
    if (g_al.hist_cost) g_al.hist_cost[static_cast<size_t>(rep) * batch + b] = scost[b];
    if (g_al.hist_iters) g_al.hist_iters[static_cast<size_t>(rep) * batch + b] = sit[b];
  }
 }
  for (int b = 0; b < batch; ++b) {
    for (int r = 0; r < NX * (T + 1); ++r) X[static_cast<size_t>(b) * NX * (T + 1) + r] = sX[static_cast<size_t>(r) * ld + b];
    for (int r = 0; r < NU * T; ++r) U[static_cast<size_t>(b) * NU * T + r] = sU[static_cast<size_t>(r) * ld + b];
    cost[b] = scost[b];
    iters[b] = sit[b];
    status[b] = sst[b];
    if (trials) trials[b] = str[b];
    if (reg) reg[b] = srg[b];
  }
  return 0;
}

}  // namespace

extern "C" int emu_ilqr_solve_batch(int model, int batch, int T, double dt, unsigned mask, int has_bounds, const double* lo, const double* hi,
                                    const double* shared_p, const double* per_problem_p, const double* x0, double* U, double* X, double* cost,
                                    int* iters, int* status, int* trials, int* reg, int max_iterations, double tolerance, int L, int C) {
  switch (model) {
    case 0: return emulate<StLane>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
    case 1: return emulate<StCirc>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
    case 2: return emulate<Lqr4>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
    case 3: return emulate<Pendulum>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
    case 4: return emulate<Rocket>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
    case 5: return emulate<StLaneCon>(batch, T, dt, mask, has_bounds, lo, hi, shared_p, per_problem_p, x0, U, X, cost, iters, status, trials, reg, max_iterations, tolerance, L, C);
  }
  return 1;
}

// Settings for the next emu_ilqr_solve_batch calls on constrained models (pass repeats = 1 and nulls to reset).
extern "C" void emu_set_trial_store(int enable) { g_al.trial_store = enable != 0; }
extern "C" void emu_set_backward_lanes(int lanes) { g_al.backward_lanes = lanes; }
extern "C" void emu_set_sweep_lanes(int on) { g_al.sweep_lanes = on != 0; }
extern "C" void emu_set_sweep_wide(int mode) { g_al.sweep_wide = mode; }

extern "C" void emu_set_al_options(double penalty, double penalty_increase, double constraint_tolerance, double activation_tolerance, int repeats,
                                   double* hist_cost, int* hist_iters) {
  g_al.penalty = penalty;
  g_al.penalty_increase = penalty_increase;
  g_al.constraint_tolerance = constraint_tolerance;
  g_al.activation_tolerance = activation_tolerance;
  g_al.repeats = repeats < 1 ? 1 : repeats;
  g_al.hist_cost = hist_cost;
  g_al.hist_iters = hist_iters;
}

// ---- centralized (stacked) solve, run with tid = 0, nthr = 1 -----------------------------------------------
#include "centralized.cuh"

namespace {
template <class M>
int emulate_centralized(int A, int T, double dt, int has_bounds, const double* lo, const double* hi, const double* params_per_agent,
                        const double* x0, double* U /* [T][ms] in/out */, double* X /* [T+1][ns] */, double* out_cost, int* out_int,
                        int max_iterations, double tolerance) {
  constexpr int NPs = (M::NP > 0 ? M::NP : 1);
  const StackedWork W(A, M::NX, M::NU);
  const int ns = W.ns, ms = W.ms;
  std::vector<double> Xt(static_cast<size_t>(T + 1) * ns), Ut(static_cast<size_t>(T) * ms), K(static_cast<size_t>(T) * ms * ns), k(static_cast<size_t>(T) * ms),
      work(W.total, 0.0), prm(static_cast<size_t>(A) * NPs, 0.0);
  for (int a = 0; a < A; ++a)
    for (int i = 0; i < M::NP; ++i) prm[a * NPs + i] = params_per_agent[a * M::NP + i];
  StackedProblem<M> P{};
  P.A = A;
  P.T = T;
  P.dt = dt;
  P.has_bounds = has_bounds;
  for (int i = 0; i < M::NU; ++i) {
    P.lo[i] = lo[i];
    P.hi[i] = hi[i];
  }
  P.tolerance = tolerance;
  P.max_iterations = max_iterations;
  P.x0 = x0;
  P.prm = prm.data();
  P.X = X;
  P.U = U;
  P.Xt = Xt.data();
  P.Ut = Ut.data();
  P.K = K.data();
  P.kff = k.data();
  P.work = work.data();
  P.fast = work.data() + W.fast;
  P.out_cost = out_cost;
  P.out_int = out_int;
  P.phase_cycles = nullptr;
  P.use_dmma = 0;
  P.max_ms = std::numeric_limits<double>::infinity();
  stacked_solve<M>(P, 0, 1);
  return 0;
}
}  // namespace

extern "C" int emu_centralized_solve(int model, int A, int T, double dt, int has_bounds, const double* lo, const double* hi,
                                     const double* params_per_agent, const double* x0, double* U, double* X, double* out_cost, int* out_int,
                                     int max_iterations, double tolerance) {
  switch (model) {
    case 1: return emulate_centralized<StCirc>(A, T, dt, has_bounds, lo, hi, params_per_agent, x0, U, X, out_cost, out_int, max_iterations, tolerance);
    case 2: return emulate_centralized<Lqr4>(A, T, dt, has_bounds, lo, hi, params_per_agent, x0, U, X, out_cost, out_int, max_iterations, tolerance);
  }
  return 1;
}

// ---- centralized strategy over agents of different models: stacked_mixed.cuh with tid = 0, nthr = 1 --------------------------
#include "stacked_mixed.cuh"

extern "C" int emu_centralized_mixed(int n_agents, const int* models, int T, double dt, int has_bounds, const double* lo, const double* hi,
                                     const double* params /* [n_agents][kMaxParams] */, const double* x0, double* X, double* U, double* out_cost,
                                     int* out_int, int max_iterations, double tolerance) {
  std::vector<MixedBlock> blocks(n_agents);
  int ns = 0, ms = 0;
  for (int a = 0; a < n_agents; ++a) {
    MixedBlock& b = blocks[a];
    b.model_id = models[a];
    if (!mixed_model_dims(b.model_id, &b.nx, &b.nu)) return 1;
    b.state_offset = ns;
    b.control_offset = ms;
    ns += b.nx;
    ms += b.nu;
    for (int i = 0; i < kMaxParams; ++i) b.params[i] = 0.0;
  }
  std::vector<int> box(ns), bou(ms);
  for (int a = 0; a < n_agents; ++a) {
    for (int i = 0; i < blocks[a].nx; ++i) box[blocks[a].state_offset + i] = a;
    for (int i = 0; i < blocks[a].nu; ++i) bou[blocks[a].control_offset + i] = a;
  }
  const MixedWork W(ns, ms, T, n_agents);
  std::vector<double> work(W.total, 0.0);
  MixedStacked P{};
  P.n_blocks = n_agents;
  P.ns = ns;
  P.ms = ms;
  P.T = T;
  P.dt = dt;
  P.has_bounds = has_bounds;
  P.tolerance = tolerance;
  P.max_iterations = max_iterations;
  P.max_ms = std::numeric_limits<double>::infinity();
  P.blocks = blocks.data();
  P.block_of_x = box.data();
  P.block_of_u = bou.data();
  P.lo = lo;
  P.hi = hi;
  P.x0 = x0;
  P.prm = params;
  P.X = X;
  P.U = U;
  P.work = work.data();
  P.out_cost = out_cost;
  P.out_int = out_int;
  mixed_stacked_solve(P, 0, 1);
  return 0;
}
