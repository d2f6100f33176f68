// centralized_threads_test.cpp -- TEST HARNESS: the barrier structure of centralized.cuh (config 5's kernel) under real threads.
//
// stacked_solve<M> runs data-parallel phases between CTA-wide barriers and, from 128 threads up, overlaps the Q_uu
// factorisation (threads 0..63, named barrier 1) with the finite differences of the next time step (the other threads,
// named barrier 2).  The parity tests run this source with one thread, which cannot see a missing or mis-sized barrier.
// Here NTHR host threads play the threads of the CTA: MAS_HOST_THREADS_SYNC maps the CTA barrier and the two named
// barriers to pthread barriers of the right participant counts.
//   * built with -fsanitize=thread, ThreadSanitizer reports conflicting accesses that no barrier orders;
//   * results must equal the one-thread run bit for bit for every thread count (with and without the overlap).
// CENTRALIZED_DROP_BARRIER=k: every thread skips its k-th CTA-wide barrier -- the self-test of the detector.
// Built and run by tests/test_host_emulation.py::test_centralized_kernel_has_no_races_between_barriers.
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

static pthread_barrier_t g_cta, g_named[3];
static int g_named_count[3] = {0, 0, 0};
static bool g_threaded = false;
static long g_drop = -1;
static thread_local long t_barriers = 0;
static long g_named_waits = 0;  // by thread 0 of group 1: shows that the overlapped path ran
static inline void host_threads_sync(int id, int count, int all) {
  (void)all;
  if (!g_threaded) return;
  if (id == 0) {
    if (++t_barriers == g_drop) return;
    pthread_barrier_wait(&g_cta);
  } else {
    if (count != g_named_count[id]) {
      std::fprintf(stderr, "named barrier %d used with %d participants, expected %d\n", id, count, g_named_count[id]);
      std::abort();
    }
    if (pthread_barrier_wait(&g_named[id]) == PTHREAD_BARRIER_SERIAL_THREAD && id == 1) ++g_named_waits;
  }
}
#define MAS_HOST_THREADS_SYNC(id, count, all) host_threads_sync(id, count, all)

#include "centralized.cuh"

using namespace mas_b200;

namespace {

struct Result {
  std::vector<double> X, U, cost;
  int ints[4];
  bool operator==(const Result& o) const {
    return std::memcmp(X.data(), o.X.data(), X.size() * 8) == 0 && std::memcmp(U.data(), o.U.data(), U.size() * 8) == 0 &&
           std::memcmp(cost.data(), o.cost.data(), cost.size() * 8) == 0 && std::memcmp(ints, o.ints, sizeof(ints)) == 0;
  }
};

// A circular-track agents (multi_agent_single_track.cpp:36-44), stacked: config 5 at a small agent count
Result run(int A, int T, int max_iterations, int nthr) {
  using M = StCirc;
  constexpr int NPs = M::NP;
  const StackedWork W(A, M::NX, M::NU);
  const int ns = W.ns, ms = W.ms;
  std::vector<double> Xt(static_cast<size_t>(T + 1) * ns), Ut(static_cast<size_t>(T) * ms), K(static_cast<size_t>(T) * ms * ns), k(static_cast<size_t>(T) * ms),
      work(W.total, 0.0), prm(static_cast<size_t>(A) * NPs, 0.0), x0(ns);
  const double defaults[6] = {20.0, 5.0, 1.0, 1.0, 0.001, 0.001};
  for (int a = 0; a < A; ++a) {
    for (int i = 0; i < M::NP; ++i) prm[a * NPs + i] = defaults[i];
    const double th = 2.0 * M_PI * a / A;
    x0[a * 4 + 0] = 20.0 * std::cos(th);
    x0[a * 4 + 1] = 20.0 * std::sin(th);
    x0[a * 4 + 2] = 1.57 + th;
    x0[a * 4 + 3] = 4.0;
  }
  Result res;
  res.X.assign(static_cast<size_t>(T + 1) * ns, 0.0);
  res.U.assign(static_cast<size_t>(T) * ms, 0.0);
  res.cost.assign(1 + A, 0.0);
  StackedProblem<M> P{};
  P.A = A;
  P.T = T;
  P.dt = 0.5;
  P.has_bounds = 1;
  for (int i = 0; i < M::NU; ++i) {
    P.lo[i] = -0.5;
    P.hi[i] = 0.5;
  }
  P.tolerance = 1e-5;
  P.max_iterations = max_iterations;
  P.x0 = x0.data();
  P.prm = prm.data();
  P.X = res.X.data();
  P.U = res.U.data();
  P.Xt = Xt.data();
  P.Ut = Ut.data();
  P.K = K.data();
  P.kff = k.data();
  P.work = work.data();
  P.fast = work.data() + W.fast;
  P.out_cost = res.cost.data();
  P.out_int = res.ints;
  P.phase_cycles = nullptr;
  P.use_dmma = 0;
  P.max_ms = std::numeric_limits<double>::infinity();
  if (nthr == 1) {
    g_threaded = false;
    stacked_solve<M>(P, 0, 1);
  } else {
    g_threaded = true;
    pthread_barrier_init(&g_cta, nullptr, nthr);
    const bool overlap = nthr >= 128 && nthr % 32 == 0;  // stacked_backward: threads 0..63 factorise, the rest differentiate
    g_named_count[1] = overlap ? 64 : 0;
    g_named_count[2] = overlap ? nthr - 64 : 0;
    if (overlap) {
      pthread_barrier_init(&g_named[1], nullptr, 64);
      pthread_barrier_init(&g_named[2], nullptr, nthr - 64);
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; ++t)
      th.emplace_back([&P, t, nthr] {
        t_barriers = 0;
        stacked_solve<M>(P, t, nthr);
      });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&g_cta);
    if (overlap) {
      pthread_barrier_destroy(&g_named[1]);
      pthread_barrier_destroy(&g_named[2]);
    }
    g_threaded = false;
  }
  return res;
}

}  // namespace

int main(int argc, char** argv) {
  if (const char* d = std::getenv("CENTRALIZED_DROP_BARRIER")) g_drop = std::atol(d);
  const int A = argc > 1 ? std::atoi(argv[1]) : 3, T = argc > 2 ? std::atoi(argv[2]) : 4, iters = argc > 3 ? std::atoi(argv[3]) : 3;
  int failures = 0;
  const Result one = run(A, T, iters, 1);
  std::printf("%d agents, T %d: one thread: iterations %d, retries %d, candidates %d, cost %.17g\n", A, T, one.ints[0], one.ints[2], one.ints[3], one.cost[0]);
  if (one.ints[0] < 1) ++failures;
  for (int nthr : {7, 32, 128, 160}) {
    const Result many = run(A, T, iters, nthr);
    const bool same = many == one;
    std::printf("  %3d threads%s: %s\n", nthr, (nthr >= 128 && nthr % 32 == 0) ? " (factorisation overlapped with the next step's finite differences)" : "",
                same ? "identical to one thread" : "DIFFERS");
    if (nthr >= 128 && nthr % 32 == 0 && g_named_waits == 0) {
      std::printf("  the overlapped path did not run\n");
      ++failures;
    }
    failures += same ? 0 : 1;
  }
  if (failures) return 1;
  std::printf("ALL OK\n");
  return 0;
}
