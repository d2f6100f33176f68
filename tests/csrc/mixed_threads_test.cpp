// mixed_threads_test.cpp -- TEST HARNESS: the barrier structure of stacked_mixed.cuh under real threads.
//
// The general stacked solve (centralized strategy over agents of different models) is written as data-parallel phases
// `for (idx = tid; idx < n; idx += nthr)` between MAS_CTA_SYNC() barriers.  The host emulation used by the parity tests runs
// it with one thread, which cannot see a missing barrier.  Here MAS_CTA_SYNC() is a pthread barrier and the same source runs
// with NTHR host threads playing the threads of a CTA:
//   * built with -fsanitize=thread, ThreadSanitizer reports any pair of accesses to the workspace / result arrays that
//     the barriers do not order (a missing or misplaced __syncthreads() on the GPU);
//   * the results must equal the one-thread run bit for bit, whatever the thread count.
// Built and run by tests/test_host_emulation.py::test_mixed_stacked_solve_has_no_races_between_barriers.
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

static pthread_barrier_t g_barrier;
static bool g_threaded = false;
// Self-test of the harness: with MIXED_DROP_BARRIER=k in the environment every thread skips its k-th barrier (all threads
// execute the same barrier sequence, so nothing deadlocks) -- the detector must then report the race it uncovers.
static long g_drop = -1;
static thread_local long t_barriers = 0;
static inline void host_cta_sync() {
  if (!g_threaded) return;
  if (++t_barriers == g_drop) return;
  pthread_barrier_wait(&g_barrier);
}
#define MAS_HOST_THREADS_SYNC(id, count, all) host_cta_sync()  // centralized.cuh: MAS_CTA_SYNC() on the host

#include "stacked_mixed.cuh"

using namespace mas_b200;

namespace {

struct ModelRow {
  int T;
  double dt;
  int has_bounds;
  double lo[4], hi[4], prm[8];
  double x0[4];
};
// the example OCPs (tests/conftest.py MODEL_TABLE) with one fixed initial state each
const ModelRow kRows[5] = {
    {80, 0.1, 1, {-0.7, -1.0}, {0.7, 1.0}, {1.0, 10.0, 1.0, 0.1, 0.1}, {0.0, 1.2, -0.2, 0.7}},
    {10, 0.5, 1, {-0.5, -0.5}, {0.5, 0.5}, {20.0, 5.0, 1.0, 1.0, 0.001, 0.001}, {14.0, 14.3, 2.4, 4.0}},
    {10, 0.1, 0, {0, 0, 0, 0}, {0, 0, 0, 0}, {0}, {0.4, -0.7, 0.1, 0.9}},
    {60, 0.05, 1, {-5.0}, {5.0}, {60.0}, {3.05, 0.07}},
    {50, 0.1, 1, {0.0}, {20.0}, {9.81, 50.0, 5e-3, 15.0, 2.0, 0.0}, {0.3, -0.4, 1.1}},
};

struct Result {
  std::vector<double> X, U, cost;
  int ints[4];
  bool operator==(const Result& o) const {
    return X.size() == o.X.size() && U.size() == o.U.size() && std::memcmp(X.data(), o.X.data(), X.size() * 8) == 0 &&
           std::memcmp(U.data(), o.U.data(), U.size() * 8) == 0 && std::memcmp(cost.data(), o.cost.data(), cost.size() * 8) == 0 &&
           std::memcmp(ints, o.ints, sizeof(ints)) == 0;
  }
};

Result run(const std::vector<int>& models, int T, int max_iterations, int nthr) {
  const int A = static_cast<int>(models.size());
  std::vector<MixedBlock> blocks(A);
  int ns = 0, ms = 0, all_bounds = 1;
  for (int a = 0; a < A; ++a) {
    MixedBlock& b = blocks[a];
    b.model_id = models[a];
    mixed_model_dims(b.model_id, &b.nx, &b.nu);
    b.state_offset = ns;
    b.control_offset = ms;
    ns += b.nx;
    ms += b.nu;
    for (int i = 0; i < kMaxParams; ++i) b.params[i] = 0.0;
    all_bounds = all_bounds && kRows[models[a]].has_bounds;
  }
  std::vector<int> box(ns), bou(ms);
  std::vector<double> lo(ms), hi(ms), x0(ns), prm(static_cast<size_t>(A) * kMaxParams, 0.0);
  for (int a = 0; a < A; ++a) {
    const ModelRow& r = kRows[models[a]];
    for (int i = 0; i < blocks[a].nx; ++i) {
      box[blocks[a].state_offset + i] = a;
      x0[blocks[a].state_offset + i] = r.x0[i] + 0.01 * a;
    }
    for (int i = 0; i < blocks[a].nu; ++i) {
      bou[blocks[a].control_offset + i] = a;
      lo[blocks[a].control_offset + i] = r.lo[i];
      hi[blocks[a].control_offset + i] = r.hi[i];
    }
    for (int i = 0; i < 8 && i < kMaxParams; ++i) prm[static_cast<size_t>(a) * kMaxParams + i] = r.prm[i];
  }
  const MixedWork W(ns, ms, T, A);
  std::vector<double> work(W.total, 0.0);
  Result res;
  res.X.assign(static_cast<size_t>(T + 1) * ns, 0.0);
  res.U.assign(static_cast<size_t>(T) * ms, 0.0);
  res.cost.assign(1 + A, 0.0);
  MixedStacked P{};
  P.n_blocks = A;
  P.ns = ns;
  P.ms = ms;
  P.T = T;
  P.dt = kRows[models[0]].dt;
  P.has_bounds = all_bounds;
  P.tolerance = 1e-5;
  P.max_iterations = max_iterations;
  P.max_ms = std::numeric_limits<double>::infinity();
  P.blocks = blocks.data();
  P.block_of_x = box.data();
  P.block_of_u = bou.data();
  P.lo = lo.data();
  P.hi = hi.data();
  P.x0 = x0.data();
  P.prm = prm.data();
  P.X = res.X.data();
  P.U = res.U.data();
  P.work = work.data();
  P.out_cost = res.cost.data();
  P.out_int = res.ints;
  if (nthr == 1) {
    g_threaded = false;
    mixed_stacked_solve(P, 0, 1);
  } else {
    g_threaded = true;
    pthread_barrier_init(&g_barrier, nullptr, nthr);
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; ++t)
      th.emplace_back([&P, t, nthr] {
        t_barriers = 0;
        mixed_stacked_solve(P, t, nthr);
      });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_barrier);
    g_threaded = false;
  }
  return res;
}

}  // namespace

int main() {
  int failures = 0;
  if (const char* d = std::getenv("MIXED_DROP_BARRIER")) g_drop = std::atol(d);
  const std::vector<std::vector<int>> mixes = {{1, 0, 3, 2, 4}, {4, 3, 1}, {3, 4}, {2, 1, 2}};
  for (const auto& models : mixes) {
    const int T = std::min(kRows[models[0]].T, 12);  // horizon of the first block (build_global_ocp), shortened: the phases are the same
    const Result one = run(models, T, 4, 1);
    for (int nthr : {2, 5, 8}) {
      const Result many = run(models, T, 4, nthr);
      const bool same = many == one;
      std::printf("mix of %zu agents, T %d, %d threads: iterations %d, candidates %d, cost %.17g -> %s\n", models.size(), T, nthr, many.ints[0],
                  many.ints[3], many.cost[0], same ? "identical to one thread" : "DIFFERS");
      failures += same ? 0 : 1;
    }
    if (one.ints[0] < 1) {
      std::printf("no iteration ran\n");
      ++failures;
    }
  }
  if (failures) return 1;
  std::printf("ALL OK\n");
  return 0;
}
