// lanes_threads_test.cpp -- TEST HARNESS: the lane-cooperative backward passes of the batched engine under real threads.
//
//   * RiccatiLanes (riccati_sweep_lanes_kernel, engine.cuh): the lanes of a problem, one per column of V_xx, run phases a..e of a
//     Riccati step with a __syncwarp() between the phases and exchange their columns through shared memory.  Default backward
//     pass of small active sets (the last iterations of the headline batch, the single-solve latency path).
//   * backward_lanes (backward_lanes_kernel): eight lanes deal out the finite-difference stencil points of a step, lane 0
//     gathers them after a group barrier and runs the Riccati step.  FD-heavy derivative modes (config 2).
// The host emulation behind the parity tests runs the lanes one after the other.  Here every lane is a host thread, the
// barriers are pthread barriers PLACED AS IN THE KERNELS, and the program is built with -fsanitize=thread: the sanitizer
// reports any access pair the barriers leave unordered, and gains / retries must equal the one-thread sweep bit for bit.
// LANES_DROP_BARRIER=k: every lane skips its k-th barrier (self-test of the detector).
// Built and run by tests/test_host_emulation.py::test_lane_cooperative_backward_passes_have_no_races.
#include <pthread.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ilqr_core.cuh"

using namespace mas_b200;

namespace {

pthread_barrier_t g_barrier;
long g_drop = -1;
thread_local long t_barriers = 0;
void lane_barrier() {
  if (++t_barriers == g_drop) return;
  pthread_barrier_wait(&g_barrier);
}

// one single-track lane-following problem (config 1 / 3) in the engine's [T][dim][ld] layout, slot p of a 32-wide column
template <class M>
struct Problem {
  static constexpr int NX = M::NX, NU = M::NU, ld = 32;
  int T;
  std::vector<double> x0, X, U, K, kff, params;
  std::vector<double> cost, merit, pen;
  std::vector<int> iters, status, trials, reg;
  BatchView<NX, NU> v{};
  Problem(int T_, double dt, unsigned mask, const double* lo, const double* hi, const double* shared_p, const double* x0_, int p) : T(T_) {
    x0.assign(static_cast<size_t>(NX) * ld, 0.0);
    X.assign(static_cast<size_t>(NX) * (T + 1) * ld, 0.0);
    U.assign(static_cast<size_t>(NU) * T * ld, 0.0);
    K.assign(static_cast<size_t>(NU) * NX * T * ld, 0.0);
    kff.assign(static_cast<size_t>(NU) * T * ld, 0.0);
    params.assign(static_cast<size_t>(kMaxParams) * ld, 0.0);
    cost.assign(ld, 0.0);
    merit.assign(ld, 0.0);
    pen.assign(ld, 10.0);
    iters.assign(ld, 0);
    status.assign(ld, 0);
    trials.assign(ld, 0);
    reg.assign(ld, 0);
    for (int i = 0; i < NX; ++i) x0[static_cast<size_t>(i) * ld + p] = x0_[i];
    for (int t = 0; t < T; ++t)  // a nominal control sequence that is not zero: steer and accelerate a little
      for (int i = 0; i < NU; ++i) U[soa_index<NU>(t, i, ld, p)] = 0.05 * (i + 1) * ((t % 7) - 3) / 3.0;
    v.ld = ld;
    v.T = T;
    v.set_dt(dt);
    v.deriv_mask = mask;
    v.set_bounds(1, lo, hi);
    v.per_problem_params = 0;
    for (int i = 0; i < kMaxParams; ++i) v.shared_p[i] = i < M::NP ? shared_p[i] : 0.0;
    v.params = params.data();
    v.x0 = x0.data();
    v.X = X.data();
    v.U = U.data();
    v.K = K.data();
    v.kff = kff.data();
    v.cost = cost.data();
    v.merit = merit.data();
    v.iters = iters.data();
    v.status = status.data();
    v.trials = trials.data();
    v.reg_retries = reg.data();
    v.penalty = pen.data();
    v.tolerance = 1e-5;
    v.max_iterations = 10;
    rollout_thread<M>(v, p);
  }
};

template <class M, int MASK_CT>
int check(const char* what, unsigned mask, int T, double dt, const double* lo, const double* hi, const double* prm, const double* x0) {
  constexpr int NX = M::NX;
  using D = DerivBlock<M>;
  using RL = RiccatiLanes<M, MASK_CT>;
  const int p = 5;
  int failures = 0;
  // reference: derivative blocks of all points, then the one-thread sweep
  Problem<M> ref(T, dt, mask, lo, hi, prm, x0, p);
  std::vector<double> store(static_cast<size_t>(T + 1) * D::size, 0.0);
  for (int t = T; t >= 0; --t) linearize_point<M>(ref.v, p, t, mask, 0, 1, [&](int off, double val) { store[static_cast<size_t>(t) * D::size + off] = val; });
  const int retries_ref = riccati_sweep_thread<M, MASK_CT>(ref.v, p, [&](int t, double* blk) {
    for (int k = 0; k < D::size; ++k) blk[k] = store[static_cast<size_t>(t) * D::size + k];
  });

  double kmax = 0.0;
  for (double g : ref.K) kmax = std::fabs(g) > kmax ? std::fabs(g) : kmax;
  if (!(kmax > 0.0)) {
    std::printf("%s: the reference sweep produced no gains\n", what);
    ++failures;
  }
  // (1) RiccatiLanes: LG lanes, barriers exactly where riccati_sweep_lanes_kernel has its __syncwarp()s
  {
    Problem<M> q(T, dt, mask, lo, hi, prm, x0, p);
    constexpr int LG = RL::LG;
    std::vector<double> xch(RL::XCH, 0.0);
    std::vector<int> retries(LG, 0);
    pthread_barrier_init(&g_barrier, nullptr, LG);
    std::vector<std::thread> th;
    for (int j = 0; j < LG; ++j)
      th.emplace_back([&, j] {
        t_barriers = 0;
        const bool active = j < NX;  // lanes beyond the state dimension only take part in the barriers
        RL r;
        lane_barrier();  // the terminal block has landed (stage_wait + __syncwarp)
        if (active) r.init_terminal(&store[static_cast<size_t>(T) * D::size], j);
        for (int t = T - 1; t >= 0; --t) {
          const double* blk = &store[static_cast<size_t>(t) * D::size];
          lane_barrier();
          if (active) r.phase_a(blk, j, xch.data());
          lane_barrier();
          if (active) r.phase_b(blk, j, xch.data());
          lane_barrier();
          if (active) r.phase_c(j, xch.data());
          lane_barrier();
          if (active) r.phase_d(q.v, p, t, j, xch.data());
          lane_barrier();
          if (active) r.phase_e(j, xch.data());
        }
        retries[j] = r.retries;
      });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_barrier);
    const bool same = q.K == ref.K && q.kff == ref.kff && retries[0] == retries_ref;
    std::printf("%s: RiccatiLanes, %d lanes, T %d: retries %d -> %s\n", what, LG, T, retries[0], same ? "identical to the one-thread sweep" : "DIFFERS");
    failures += same ? 0 : 1;
  }
  // (2) backward_lanes: 8 lanes, the group barrier is the functor the kernel passes (__syncwarp over the group's lanes)
  {
    Problem<M> q(T, dt, mask, lo, hi, prm, x0, p);
    constexpr int LB = 8;
    std::vector<double> blk(D::size > D::n_terminal_tasks ? D::size : D::n_terminal_tasks, 0.0);
    std::vector<int> retries(LB, 0);
    pthread_barrier_init(&g_barrier, nullptr, LB);
    std::vector<std::thread> th;
    for (int lane = 0; lane < LB; ++lane)
      th.emplace_back([&, lane] {
        t_barriers = 0;
        retries[lane] = backward_lanes<M, MASK_CT, LB>(q.v, p, lane, blk.data(), [] { lane_barrier(); });
      });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_barrier);
    const bool same = q.K == ref.K && q.kff == ref.kff && retries[0] == retries_ref;
    std::printf("%s: backward_lanes, %d lanes, T %d: retries %d -> %s\n", what, LB, T, retries[0], same ? "identical to the one-thread sweep" : "DIFFERS");
    failures += same ? 0 : 1;
  }
  return failures;
}

}  // namespace

int main() {
  if (const char* d = std::getenv("LANES_DROP_BARRIER")) g_drop = std::atol(d);
  int failures = 0;
  {
    const double lo[2] = {-0.7, -1.0}, hi[2] = {0.7, 1.0}, prm[5] = {1.0, 10.0, 1.0, 0.1, 0.1}, x0[4] = {0.0, 1.3, -0.2, 0.6};
    failures += check<StLane, StLane::EXAMPLE_MASK>("single-track lane, example derivatives", StLane::EXAMPLE_MASK, 80, 0.1, lo, hi, prm, x0);
    failures += check<StLane, 0>("single-track lane, all finite differences", 0u, 40, 0.1, lo, hi, prm, x0);
  }
  {
    const double lo[2] = {-0.5, -0.5}, hi[2] = {0.5, 0.5}, prm[6] = {20.0, 5.0, 1.0, 1.0, 0.001, 0.001}, x0[4] = {14.0, 14.3, 2.4, 4.0};
    failures += check<StCirc, 0>("single-track circle, all finite differences", 0u, 10, 0.5, lo, hi, prm, x0);
  }
  if (failures) return 1;
  std::printf("ALL OK\n");
  return 0;
}
