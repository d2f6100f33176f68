// engine.cuh -- kernels and the host-side batch engine behind the C ABI (include/mas_b200.h).
//
// One Batch = `batch` same-shaped OCPs resident in HBM in the [T][dim][ld] layout of ilqr_core.cuh.
// An iLQR solve of the whole batch (mas::solve(Solver&, OCP&) for every problem,
// solvers/solver.hpp:28-32 -> solvers/ilqr.hpp:59-273) is a host loop over iterations that launches
//   backward_kernel      : one thread per still-active problem; derivatives + Riccati + gains (ilqr.hpp:92-193)
//   forward_coop_kernel  : the line search (ilqr.hpp:195-271) for large active sets: a warp owns 32 problems and
//                          deals its lanes out to the (problem, step size) rollouts still needed
//   forward_kernel       : the line search for small active sets: L lanes per problem roll out all step sizes
//                          concurrently, warp-shuffle selection of the first improving one
//   both then take the accepted step (re-rolled in place, or copied from the trial store in the wide lane mappings),
//   apply the stop test and compact the active list.
// Problems leave the active list as they converge, so later iterations only pay for what is left.
// The strategy layer's kernels (trust region, joint line search) and the timing hooks are here too; the
// centralized strategy lives in centralized.cuh.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ilqr_core.cuh"
#include "mas_b200.h"

namespace mas_b200 {

void set_last_error(const std::string& msg);

#define MAS_CUDA_CHECK(expr)                                                                              \
  do {                                                                                                    \
    cudaError_t err__ = (expr);                                                                           \
    if (err__ != cudaSuccess) {                                                                           \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(err__));                              \
      return MAS_B200_ERR_CUDA;                                                                           \
    }                                                                                                     \
  } while (0)

// 64-thread CTAs: the iLQR kernels need 120-250 registers per thread, and with 2-warp CTAs the
// register file of an SM (64K) divides into more resident CTAs than with 4-warp ones (e.g. 7 x 64
// instead of 3 x 128 threads at 134 registers), which turns the 65,536-problem batch into one wave.
constexpr int kBlock = 64;
// minimum resident CTAs per SM asked of the compiler for the iLQR kernels (7 -> 128 registers, in practice 8 CTAs)
// launches of at most this many threads (under a quarter of the device) stage their operands in shared memory
constexpr int kStageMaxThreads = 16384;
#ifndef MAS_MIN_CTAS
#define MAS_MIN_CTAS 7
#endif

struct Context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  bool blocking_sync = false;  // host waits on events sleep instead of spinning (many pipelines per core: mas_b200_context_set_blocking_sync)
  void* nccl_comm = nullptr;  // ncclComm_t when multi-GPU is initialised
  int rank = 0, world = 1;
  // device workspace of the last centralized strategy call, kept for the next call of the same shape (a 592-scenario
  // config-5 run needs 1.3 GB: allocating and freeing it per call cost up to 100 ms); freed with the context
  std::shared_ptr<void> centralized_workspace;
  long long centralized_key[4] = {0, 0, 0, 0};
};

// ---- kernels ------------------------------------------------------------------------------------
// Batched transposes between the caller's [batch][rows] arrays and the [rows][ld] HBM layout.
__global__ void aos_to_soa_kernel(const double* __restrict__ src, double* __restrict__ dst, int batch, int rows, int ld);
__global__ void soa_to_aos_kernel(const double* __restrict__ src, double* __restrict__ dst, int batch, int rows, int ld);
__global__ void dfma_probe_kernel(double* out, int iters);
__global__ void division_selftest_kernel(unsigned long long seed, int per_thread, unsigned long long* out);
__global__ void publish_count_kernel(const int* __restrict__ count, int* mapped_host_word);

// OCP::initialize_problem / iLQR prologue: rollout + cost, reset of the per-solve counters and of
// the active list (identity).  ilqr.hpp:75-78, ocp.hpp:110-113,182.
template <class M>
__global__ void __launch_bounds__(kBlock) prologue_kernel(BatchView<M::NX, M::NU> v, int batch, int* list, int* count, int max_iterations) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) *count = (max_iterations > 0) ? batch : 0;
  if (p >= batch) return;
  const double c = rollout_thread<M>(v, p);
  v.cost[p] = c;
  if (HasConstraints<M>::value) {
    double prm[M::NP > 0 ? M::NP : 1];
    load_params<M>(v, p, prm);
    v.merit[p] = al_merit_of_stored<M>(v, p, prm, c);  // compute_merit (ilqr.hpp:78,380-407) with the current multipliers
  } else {
    v.merit[p] = c;  // compute_merit == objective without constraint callbacks
  }
  v.iters[p] = 0;
  v.trials[p] = 0;
  v.reg_retries[p] = 0;
  v.status[p] = STATUS_MAX_ITER;
  list[p] = p;
  if (v.dbg && v.dbg_records > 0) {  // "iLQR initial cost=... merit=..." (ilqr.hpp:79-80)
    v.dbg[p] = c;
    v.dbg[static_cast<size_t>(v.ld) + p] = v.merit[p];
  }
}

template <class M, int MASK_CT>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) backward_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                       int* next_count) {
  extern __shared__ double s_stage[];  // dynamic: 2 * (NX + NU) * kStageStride doubles for small (latency-bound) launches
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *next_count = 0;
  if (i >= *count) return;
  const int p = list[i];
  double* stage = static_cast<int>(gridDim.x * blockDim.x) <= kStageMaxThreads ? s_stage + threadIdx.x : nullptr;
  const int r = backward_thread<M, MASK_CT>(v, p, stage);
  if (r) v.reg_retries[p] += r;
}

// Backward pass with LB lanes per problem (backward_lanes): for derivative modes with many finite-difference
// callbacks and active sets small enough for all lanes to be resident.  Launch with
// (kBlock / LB) * DerivBlock<M>::size doubles of dynamic shared memory.
struct GroupSync {
  unsigned mask;
  __device__ void operator()() const { __syncwarp(mask); }
};
template <class M, int MASK_CT, int LB>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) backward_lanes_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list,
                                                                             const int* __restrict__ count, int* next_count) {
  extern __shared__ double s_blk[];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid == 0) *next_count = 0;
  const int i = gid / LB, lane = gid % LB;
  if (i >= *count) return;  // whole groups leave together: the group barrier below only names the group's lanes
  const int p = list[i];
  const unsigned first = (threadIdx.x & 31u) & ~static_cast<unsigned>(LB - 1);
  const GroupSync sync{(LB == 32 ? 0xffffffffu : ((1u << LB) - 1u)) << first};
  double* blk = s_blk + static_cast<size_t>(threadIdx.x / LB) * DerivBlock<M>::size;
  const int r = backward_lanes<M, MASK_CT, LB>(v, p, lane, blk, sync);
  if (lane == 0 && r) v.reg_retries[p] += r;
}

// ---- time-parallel linearisation + Riccati sweep (ilqr_core.cuh: linearize_point / riccati_sweep_thread) ----------
// Small active sets are bound by the latency of the T sequential backward steps; the derivative evaluations inside a
// step do not depend on the value function, so they are taken out of that chain: linearize_kernel computes the
// derivative blocks of ALL (problem, time step) pairs at once (thread -> (problem, FD task group, t); consecutive
// threads = consecutive problems, so block entries are written as coalesced rows of D[t][entry][slot]), and
// riccati_sweep_kernel then runs the recursion alone, one thread per problem, fetching block t-1 into shared memory
// with cp.async while step t computes.  FD-heavy derivative modes gain the whole stencil (118 cost + 12 dynamics
// evaluations per step at n = 4, m = 2) spread over T x G threads per problem.
constexpr int kExportBlock = 256, kExportCtas = 16;  // result-sink kernel: PCIe-bound, a few CTAs suffice
constexpr int kLinBlock = 128;
#ifndef MAS_LIN_MIN_CTAS
#define MAS_LIN_MIN_CTAS 1
#endif
constexpr int kSweepBlock = 32;
template <class M, int MASK_CT>
__global__ void __launch_bounds__(kLinBlock, MAS_LIN_MIN_CTAS) linearize_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                             double* __restrict__ D, int cap, int n_pad, int G) {
  using DB = DerivBlock<M>;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int i = static_cast<int>(gid % n_pad);
  const long long rest = gid / n_pad;
  const int g = static_cast<int>(rest % G), t = static_cast<int>(rest / G);
  if (i >= *count || t > v.T) return;
  const int p = list[i];
  const unsigned mask = (MASK_CT >= 0) ? static_cast<unsigned>(MASK_CT) : v.deriv_mask;
  double* base = D + static_cast<size_t>(t) * DB::size * cap + i;
  linearize_point<M>(v, p, t, mask, g, G, [&](int off, double val) { base[static_cast<size_t>(off) * cap] = val; });
}

template <class M, int MASK_CT>
__global__ void __launch_bounds__(kSweepBlock) riccati_sweep_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                                   int* next_count, const double* __restrict__ D, int cap) {
  using DB = DerivBlock<M>;
  __shared__ double s_blk[2 * DB::size * kSweepBlock];  // [buffer][entry][thread]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *next_count = 0;
  if (i >= *count) return;
  const int p = list[i];
  double* mine = s_blk + threadIdx.x;
  auto issue = [&](int t) {
    const double* src = D + static_cast<size_t>(t) * DB::size * cap + i;
    double* dst = mine + static_cast<size_t>(t & 1) * DB::size * kSweepBlock;
    const int n = t == v.T ? DB::n_terminal_tasks : DB::size;
    for (int k = 0; k < n; ++k) stage_copy8(dst + k * kSweepBlock, src + static_cast<size_t>(k) * cap);
    stage_commit();
  };
  issue(v.T);
  const int r = riccati_sweep_thread<M, MASK_CT>(v, p, [&](int t, double* blk) {
    stage_wait();
    const double* src = mine + static_cast<size_t>(t & 1) * DB::size * kSweepBlock;
    const int n = t == v.T ? DB::n_terminal_tasks : DB::size;
#pragma unroll
    for (int k = 0; k < DB::size; ++k)
      if (k < n) blk[k] = src[k * kSweepBlock];
    if (t > 0) issue(t - 1);
  });
  if (r) v.reg_retries[p] += r;
}

// The Riccati recursion with the lanes of a problem sharing a step (RiccatiLanes, ilqr_core.cuh): LG lanes per problem,
// 32 / LG problems per warp, one warp per CTA.  The derivative block of step t-1 is staged into shared memory by the whole
// warp (coalesced over the warp's problems) while step t computes.
template <class M, int MASK_CT>
__global__ void __launch_bounds__(32) riccati_sweep_lanes_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                                 int* next_count, const double* __restrict__ D, int cap) {
  using DB = DerivBlock<M>;
  using RL = RiccatiLanes<M, MASK_CT>;
  constexpr int LG = RL::LG, PW = 32 / LG;  // problems per warp
  __shared__ double s_blk[2][PW][DB::size];
  __shared__ double s_xch[PW][RL::XCH];
  const int lane = threadIdx.x, q = lane / LG, j = lane % LG;
  const int first = blockIdx.x * PW;
  if (first == 0 && lane == 0) *next_count = 0;
  const int n = *count;
  if (first >= n) return;
  const int i = first + q;
  const bool active = i < n && j < M::NX;
  const int p = i < n ? list[i] : 0;
  // staging: lane -> (problem lane % PW, entries lane / PW, + LG, ...): PW consecutive slots of one entry per group of lanes
  auto issue = [&](int t) {
    const int sq = lane % PW, e0 = lane / PW;
    const int ne = t == v.T ? DB::n_terminal_tasks : DB::size;
    if (first + sq < n) {
      const double* src = D + static_cast<size_t>(t) * DB::size * cap + (first + sq);
      for (int e = e0; e < ne; e += LG) stage_copy8(&s_blk[t & 1][sq][e], src + static_cast<size_t>(e) * cap);
    }
    stage_commit();
  };
  RL r;
  issue(v.T);
  stage_wait();
  __syncwarp();
  if (v.T > 0) issue(v.T - 1);
  if (active) r.init_terminal(s_blk[v.T & 1][q], j);
  for (int t = v.T - 1; t >= 0; --t) {
    stage_wait();
    __syncwarp();  // block t is complete for every lane; nobody still reads block t+1's buffer... (it is the other buffer)
    const double* blk = s_blk[t & 1][q];
    if (active) r.phase_a(blk, j, s_xch[q]);
    __syncwarp();
    if (active) r.phase_b(blk, j, s_xch[q]);
    __syncwarp();
    if (active) r.phase_c(j, s_xch[q]);
    __syncwarp();
    if (active) r.phase_d(v, p, t, j, s_xch[q]);
    __syncwarp();
    if (t > 0) issue(t - 1);  // into the buffer block t+1 used: every lane is past its last read of it
    if (active) r.phase_e(j, s_xch[q]);
  }
  if (active && j == 0 && r.retries) v.reg_retries[p] += r.retries;
}

// The Riccati recursion with one lane per entry of the NX x NX matrices (RiccatiWide, ilqr_core.cuh): NX*NX lanes per
// problem, 32 / (NX*NX) problems per warp, one warp per CTA -- the mapping for the smallest active sets, where a launch
// costs T times the latency of one step and nothing else.  Staging of the derivative blocks as in the kernel above.
template <class M, int MASK_CT>
__global__ void __launch_bounds__(32) riccati_sweep_wide_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                                int* next_count, const double* __restrict__ D, int cap) {
  using DB = DerivBlock<M>;
  using RW = RiccatiWide<M, MASK_CT>;
  constexpr int LW = RW::LW, PW = 32 / LW;  // problems per warp
  __shared__ double s_blk[2][PW][DB::size];
  __shared__ double s_xch[PW][RW::XCH];
  const int lane = threadIdx.x, q = lane / LW, e = lane % LW;
  const int first = blockIdx.x * PW;
  if (first == 0 && lane == 0) *next_count = 0;
  const int n = *count;
  if (first >= n) return;
  const int i = first + q;
  const bool active = i < n;
  const int p = active ? list[i] : 0;
  auto issue = [&](int t) {
    const int sq = lane % PW, e0 = lane / PW;
    const int ne = t == v.T ? DB::n_terminal_tasks : DB::size;
    if (first + sq < n) {
      const double* src = D + static_cast<size_t>(t) * DB::size * cap + (first + sq);
      for (int k = e0; k < ne; k += LW) stage_copy8(&s_blk[t & 1][sq][k], src + static_cast<size_t>(k) * cap);
    }
    stage_commit();
  };
  RW r;
  double* xch = s_xch[q];
  issue(v.T);
  stage_wait();
  __syncwarp();
  if (v.T > 0) issue(v.T - 1);
  if (active) RW::init_terminal(s_blk[v.T & 1][q], e, xch);
  __syncwarp();
  if (active) RW::phase6(e, xch);
  for (int t = v.T - 1; t >= 0; --t) {
    stage_wait();
    __syncwarp();  // block t has landed for every lane, and the symmetrised V_xx of step t+1 is complete
    const double* blk = s_blk[t & 1][q];
    if (active) r.phase1(blk, e, xch);
    __syncwarp();
    if (active) r.phase2(blk, e, xch);
    __syncwarp();
    if (t > 0) issue(t - 1);  // into the buffer block t+1 used; block t is not read after phase 2
    if (active) r.phase3(v, p, t, e, xch);
    __syncwarp();
    if (active) r.phase4(e, xch);
    __syncwarp();
    if (active) r.phase5(e, xch);
    __syncwarp();
    if (active) RW::phase6(e, xch);
  }
  if (active && e == 0 && r.retries) v.reg_retries[p] += r.retries;
}

// L lanes per problem (a power of two <= 16, aligned inside the warp); lane l rolls out step sizes
// l, l+L, ...; the group picks the first improving candidate with shuffles; lane 0 commits.
template <class M, int L, int C>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) forward_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                      int* next_list, int* next_count) {
  static_assert(kBlock == kStageStride, "the staging layout assumes kBlock threads per CTA");
  // Wide mappings (small active sets, latency-bound launches) stage the next step's operands in shared memory with
  // cp.async: -19 % on those launches.  Saturating launches do not: there the extra LDGSTS + LDS traffic through the
  // memory-instruction queue costs more than the exposed load latency that other warps cover anyway (measured).
  constexpr bool kStage = L >= 4;
  extern __shared__ double s_stage[];  // dynamic: 2 * NV * kStageStride doubles when the launch stages, nothing otherwise
  double* stage = (kStage && static_cast<int>(gridDim.x * blockDim.x) <= kStageMaxThreads) ? s_stage + threadIdx.x : nullptr;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gid / L;
  const int lane = gid % L;
  const bool valid = i < *count;
  const int p = valid ? list[i] : 0;
  bool again = false;
  double prm[M::NP > 0 ? M::NP : 1];
  int best_j = kNumAlphas, best_slot = -1;
  double best_merit = 0.0, current_merit = 0.0, best_objective = 0.0;
  // Wide mappings (L >= 4: small active sets, where a launch costs the latency of its sequential passes and nothing
  // else) keep their trial trajectories: chain c of thread gid -> slot c * (threads of the grid) + gid, and the
  // accepted step becomes a copy instead of one more pass over T dependent steps.  Large active sets do not: there
  // the extra 48 B per step and rollout cost more HBM time than the second rollout (measured, 65,536 problems).
  const int n_threads = gridDim.x * blockDim.x;
  const bool store = L >= 4 && v.trial_X != nullptr && static_cast<long long>(n_threads) * C <= v.trial_slots;
  if (valid) {
    load_params<M>(v, p, prm);
    current_merit = v.merit[p];
    if (L >= 4 && store)
      lane_line_search<M, L, C, (L >= 4)>(v, p, prm, lane, current_merit, &best_j, &best_merit, gid, n_threads, &best_slot, &best_objective, stage);
    else
      lane_line_search<M, L, C, false>(v, p, prm, lane, current_merit, &best_j, &best_merit, -1, 0, nullptr, nullptr, stage);
  }
  if (L > 1) {
    // first improving candidate of the group = minimum index; its merit travels with it.
    // Executed by every lane of the warp (idle groups carry kNumAlphas) so the full mask is legal.
#pragma unroll
    for (int off = L / 2; off > 0; off >>= 1) {
      const int oj = __shfl_xor_sync(0xffffffffu, best_j, off, L);
      const double om = __shfl_xor_sync(0xffffffffu, best_merit, off, L);
      const int os = __shfl_xor_sync(0xffffffffu, best_slot, off, L);
      const double oo = __shfl_xor_sync(0xffffffffu, best_objective, off, L);
      if (oj < best_j) {
        best_j = oj;
        best_merit = om;
        best_slot = os;
        best_objective = oo;
      }
    }
  }
  if (L >= 4) __syncwarp();  // the winner's trial trajectory was written by another lane of this warp
  if (valid && lane == 0) again = finish_iteration<M>(v, p, prm, current_merit, best_j, best_merit, best_slot, best_objective);
  // warp-aggregated append to the next active list
  const unsigned vote = __ballot_sync(0xffffffffu, again);
  if (vote) {
    const int wl = threadIdx.x & 31;
    int base = 0;
    if (wl == 0) base = atomicAdd(next_count, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (again) next_list[base + __popc(vote & ((1u << wl) - 1u))] = p;
  }
}

// ---- warp-cooperative line search (large active sets) ----------------------------------------------------
// One warp owns 32 active problems.  In every round the 32 lanes are dealt out to the problems still searching
// (coop_assign): each lane rolls out one (problem, step size) task, merits go to the owner through shared memory,
// the owner takes its first improving candidate in index order -- the reference's sequential semantics
// (ilqr.hpp:206-228) -- and lanes of finished problems take over candidates of the ones that need many.  A warp's
// search costs about (candidates actually needed)/32 rollout times instead of 10.  Then owners commit / stop-test.
// STORE: lanes keep their trial trajectories (BatchView::trial_*) and an owner copies its accepted candidate right
// after the round that produced it, instead of rolling it out again at the end.
// SP: the parameters are the batch's shared ones (no per-problem table): they are read where they lie in the kernel's
// parameter block (constant-bank operands) instead of occupying registers for the whole rollout.
template <class M, int C, bool STORE, bool SP>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) forward_coop_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                                int* next_list, int* next_count) {
  constexpr int kWarps = kBlock / 32;
  __shared__ double s_merit[kWarps][32][kNumAlphas];
  __shared__ double s_obj[STORE ? kWarps : 1][STORE ? 32 : 1][kNumAlphas];  // plain cost per candidate (!= merit with constraints)
  __shared__ int s_slot[STORE ? kWarps : 1][STORE ? 32 : 1][kNumAlphas];    // trial slot holding each candidate's trajectory
  __shared__ double s_cur[kWarps][32];
  __shared__ int s_p[kWarps][32], s_done[kWarps][32], s_next[kWarps][32];
  __shared__ CoopPlan s_plan[kWarps];
  const int w = threadIdx.x / 32, lane = threadIdx.x & 31;
  const int first = (blockIdx.x * kWarps + w) * 32;
  const int n_total = *count;
  int n_valid = n_total - first;
  n_valid = n_valid < 0 ? 0 : (n_valid > 32 ? 32 : n_valid);
  const bool owner = lane < n_valid;
  const int p = owner ? list[first + lane] : 0;
  const double current_merit = owner ? v.merit[p] : 0.0;
  int accepted = -1;
  double accepted_merit = 0.0, accepted_objective = 0.0;
  bool committed = false;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x, n_threads = gridDim.x * blockDim.x;
  s_p[w][lane] = p;
  s_cur[w][lane] = current_merit;
  s_done[w][lane] = owner ? 0 : 1;
  s_next[w][lane] = 0;
  __syncwarp();
  if (n_valid > 0) {
    for (int round = 0; round < kNumAlphas; ++round) {  // at most ten rounds: every searching problem advances by >= 1
      if (lane == 0) coop_assign(s_done[w], s_next[w], n_valid, &s_plan[w], C);
      __syncwarp();
      bool any = false;
      for (int i = 0; i < n_valid; ++i) any = any || s_plan[w].quota[i] > 0;
      if (!any) break;
      const int o = s_plan[w].owner[lane];
      if (o >= 0) {
        const int po = s_p[w][o], j = s_plan[w].cand[lane];
        double prm_local[M::NP > 0 ? M::NP : 1];
        const double* prm = v.shared_p;
        if constexpr (!SP) {
          load_params<M>(v, po, prm_local);
          prm = prm_local;
        }
        double alpha[C], merit[C];
#pragma unroll
        for (int c = 0; c < C; ++c) alpha[c] = alpha_of(j + c < kNumAlphas ? j + c : kNumAlphas - 1);
        double objective[C];
        trial_rollout<M, C, STORE>(v, po, prm, alpha, merit, gid, n_threads, objective);
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (j + c < kNumAlphas) {
            s_merit[w][o][j + c] = merit[c];
            if (STORE) {
              s_obj[w][o][j + c] = objective[c];
              s_slot[w][o][j + c] = gid + c * n_threads;
            }
          }
      }
      __syncwarp();
      if (owner && !s_done[w][lane]) {
        int nx = s_next[w][lane];
        const bool fin = coop_owner_update(s_merit[w][lane], current_merit, s_plan[w].quota[lane], &nx, &accepted, &accepted_merit, C);
        s_next[w][lane] = nx;
        s_done[w][lane] = fin ? 1 : 0;
        if (STORE && accepted >= 0) {
          // take the stored trajectory now: the lane that produced it reuses its slot in the next round, and nothing
          // reads this problem's nominal trajectory any more
          accepted_objective = s_obj[w][lane][accepted];
          accept_stored<M>(v, p, s_slot[w][lane][accepted]);
          committed = true;
        }
      }
      __syncwarp();
    }
  }
  bool again = false;
  if (owner) {
    double prm_local[M::NP > 0 ? M::NP : 1];
    const double* prm = v.shared_p;
    if constexpr (!SP) {
      load_params<M>(v, p, prm_local);
      prm = prm_local;
    }
    again = finish_iteration<M>(v, p, prm, current_merit, accepted >= 0 ? accepted : kNumAlphas, accepted >= 0 ? accepted_merit : current_merit,
                                committed ? kSlotCommitted : -1, accepted_objective);
  }
  const unsigned vote = __ballot_sync(0xffffffffu, again);
  if (vote) {
    int base = 0;
    if (lane == 0) base = atomicAdd(next_count, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (again) next_list[base + __popc(vote & ((1u << lane) - 1u))] = p;
  }
}

// ---- line search as compacted rounds (large active sets) ------------------------------------------------
// Round r rolls out step sizes r*C .. r*C+C-1 for the problems that have not found an improving one
// yet; those that still have not are appended (warp-aggregated) to the next round's list.  Every warp
// of every round is full, and a problem stops costing rollouts at its first improving step size, as
// in the reference's sequential loop (ilqr.hpp:206-228).  counts[r] holds the size of round r's list.
template <class M, int C>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) trial_round_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                               int* next_list, int* next_count, int base_j, int* accept_idx,
                                                               double* accept_merit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < *count;
  const int p = valid ? list[i] : 0;
  bool keep = false;
  if (valid) {
    double prm[M::NP > 0 ? M::NP : 1];
    load_params<M>(v, p, prm);
    const double current_merit = v.merit[p];
    double alpha[C], merit[C];
#pragma unroll
    for (int c = 0; c < C; ++c) alpha[c] = alpha_of(base_j + c < kNumAlphas ? base_j + c : kNumAlphas - 1);
    trial_rollout<M, C, false>(v, p, prm, alpha, merit);
    int found = -1;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (found < 0 && base_j + c < kNumAlphas && merit[c] < current_merit) found = c;
    if (found >= 0) {
      accept_idx[p] = base_j + found;
      accept_merit[p] = merit[found];
    } else {
      keep = base_j + C < kNumAlphas;
    }
  }
  const unsigned vote = __ballot_sync(0xffffffffu, keep);
  if (vote) {
    const int wl = threadIdx.x & 31;
    int base = 0;
    if (wl == 0) base = atomicAdd(next_count, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) next_list[base + __popc(vote & ((1u << wl) - 1u))] = p;
  }
}

// After the rounds: commit the accepted step (if any), bookkeeping, stop test, next active list.
template <class M>
__global__ void __launch_bounds__(kBlock, MAS_MIN_CTAS) finish_kernel(BatchView<M::NX, M::NU> v, const int* __restrict__ list, const int* __restrict__ count,
                                                          int* next_list, int* next_count, int* accept_idx, const double* __restrict__ accept_merit) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < *count;
  const int p = valid ? list[i] : 0;
  bool again = false;
  if (valid) {
    double prm[M::NP > 0 ? M::NP : 1];
    load_params<M>(v, p, prm);
    const double current_merit = v.merit[p];
    const int aj = accept_idx[p];
    accept_idx[p] = -1;
    const int best_j = aj >= 0 ? aj : kNumAlphas;
    const double best_merit = aj >= 0 ? accept_merit[p] : current_merit;
    again = finish_iteration<M>(v, p, prm, current_merit, best_j, best_merit);
  }
  const unsigned vote = __ballot_sync(0xffffffffu, again);
  if (vote) {
    const int wl = threadIdx.x & 31;
    int base = 0;
    if (wl == 0) base = atomicAdd(next_count, __popc(vote));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (again) next_list[base + __popc(vote & ((1u << wl) - 1u))] = p;
  }
}

// Re-rollout of given controls (strategy layer, nash.hpp:224-225): X and cost from U.
template <class M>
__global__ void __launch_bounds__(kBlock) rollout_kernel(BatchView<M::NX, M::NU> v, int batch) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch) return;
  const double c = rollout_thread<M>(v, p);
  v.cost[p] = c;
}

// Trust-region accept/scale (TrustRegionNashStrategy, strategies/nash.hpp:208-243) for every agent:
// delta = U_new - U_old, Frobenius norm in column-major order; if it exceeds the radius the step is
// scaled, re-rolled out and re-costed; accept iff cand_cost < old_cost (strict) -> radius *= 1.5,
// otherwise restore the old trajectory and radius *= 0.5.
template <class M>
__global__ void __launch_bounds__(kBlock) trust_region_kernel(BatchView<M::NX, M::NU> v, int batch, const double* __restrict__ U_old,
                                                           const double* __restrict__ X_old, const double* __restrict__ cost_old, double* radius,
                                                           int* accepted) {
  constexpr int NX = M::NX, NU = M::NU;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch) return;
  const int T = v.T;
  double sq = 0.0;
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const size_t idx = soa_index<NU>(t, i, v.ld, p);
      const double d = v.U[idx] - U_old[idx];
      sq += d * d;
    }
  const double norm = sqrt(sq);
  double rad = radius[p];
  double cand_cost = v.cost[p];
  if (norm > rad) {
    const double scale = rad / norm;
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const size_t idx = soa_index<NU>(t, i, v.ld, p);
        const double d = v.U[idx] - U_old[idx];
        v.U[idx] = U_old[idx] + scale * d;
      }
    cand_cost = rollout_thread<M>(v, p);
  }
  const double oc = cost_old[p];
  if (cand_cost < oc) {
    v.cost[p] = cand_cost;
    radius[p] = rad * 1.5;
    accepted[p] = 1;
  } else {
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int i = 0; i < NU; ++i) {
        const size_t idx = soa_index<NU>(t, i, v.ld, p);
        v.U[idx] = U_old[idx];
      }
    for (int t = 0; t <= T; ++t)
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const size_t idx = soa_index<NX>(t, i, v.ld, p);
        v.X[idx] = X_old[idx];
      }
    v.cost[p] = oc;
    radius[p] = rad * 0.5;
    accepted[p] = 0;
  }
}

// ---- LineSearchNashStrategy (strategies/nash.hpp:92-180), scenario = n_agents consecutive problems -------------
// state per scenario: 0 = round accepted as solved, 1 = joint cost did not drop, searching along old -> cand,
// 2 = a trial step was accepted.  Joint costs are summed in block order from 0.0 (a fixed convention; the
// reference's OpenMP reduction order is unspecified, nash.hpp:45,134).
__global__ void nash_ls_reduce_kernel(const double* __restrict__ cost, int n_scenarios, int n_agents, double* base_cost, int* state, int phase);

// trial_controls = old + alpha * (cand - old); trial_states = rollout; trial cost (nash.hpp:136-142)
template <class M>
__global__ void __launch_bounds__(kBlock) nash_ls_trial_kernel(BatchView<M::NX, M::NU> v, int batch, int n_agents, const double* __restrict__ U_old,
                                                            const double* __restrict__ U_cand, const int* __restrict__ state, double alpha) {
  constexpr int NU = M::NU;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch || state[p / n_agents] != 1) return;
  for (int t = 0; t < v.T; ++t)
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const size_t idx = soa_index<NU>(t, i, v.ld, p);
      v.U[idx] = U_old[idx] + alpha * (U_cand[idx] - U_old[idx]);
    }
  v.cost[p] = rollout_thread<M>(v, p);
}

// no trial accepted: back to the old trajectories and costs (nash.hpp:161-171)
template <class M>
__global__ void __launch_bounds__(kBlock) nash_ls_restore_kernel(BatchView<M::NX, M::NU> v, int batch, int n_agents, const double* __restrict__ U_old,
                                                              const double* __restrict__ X_old, const double* __restrict__ cost_old,
                                                              const int* __restrict__ state) {
  constexpr int NX = M::NX, NU = M::NU;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= batch || state[p / n_agents] != 1) return;
  for (int t = 0; t < v.T; ++t)
#pragma unroll
    for (int i = 0; i < NU; ++i) {
      const size_t idx = soa_index<NU>(t, i, v.ld, p);
      v.U[idx] = U_old[idx];
    }
  for (int t = 0; t <= v.T; ++t)
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const size_t idx = soa_index<NX>(t, i, v.ld, p);
      v.X[idx] = X_old[idx];
    }
  v.cost[p] = cost_old[p];
}

// ---- host side --------------------------------------------------------------------------------------
struct BatchBase {
  Context* ctx = nullptr;
  mas_b200_ocp_desc desc{};
  int batch = 0, ld = 0, nx = 0, nu = 0, np = 0, T = 0;
  // HBM
  double *d_x0 = nullptr, *d_X = nullptr, *d_U = nullptr, *d_K = nullptr, *d_k = nullptr, *d_cost = nullptr, *d_merit = nullptr,
         *d_params = nullptr;
  int *d_iters = nullptr, *d_status = nullptr, *d_trials = nullptr, *d_reg = nullptr;
  int* d_list[2] = {nullptr, nullptr};
  int* d_count = nullptr;  // [2]
  double* d_stage = nullptr;  // staging for AoS<->SoA transposes (max(nx*(T+1), nu*T) * batch doubles)
  // asynchronous result download (begin_download / wait_download): the solution is transposed into its own staging
  // area on the solve stream, then copied to the host on a second stream while the next solve already runs
  double* d_out_stage = nullptr;  // [batch * (nx*(T+1) + nu*T + 1)] doubles, then 2 * batch ints
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_staged = nullptr, ev_downloaded = nullptr;
  bool download_pending = false;
  // trial trajectories of the line search (BatchView::trial_*), allocated at the first solve; 0 slots = recompute
  double *d_trial_X = nullptr, *d_trial_U = nullptr;
  int trial_slots = 0;
  bool trial_store = true, trial_store_tried = false;
  bool backward_lanes_enabled = true;  // lane-parallel backward pass for FD-heavy derivative modes (set_tuning lanes < 0 disables)
  // time-parallel linearisation + Riccati sweep (linearize_kernel / riccati_sweep_kernel): derivative blocks
  // D[T+1][DerivBlock::size][deriv_cap] in HBM, allocated at first use.  backward_mode: 0 auto, 1 one thread per problem
  // (fused), 2 FD tasks over eight lanes (fused), 3 time-parallel whenever the active set fits deriv_cap.
  double* d_deriv = nullptr;
  int deriv_cap = 0;
  int backward_mode = 0;
  int tp_max_problems = 8192;  // auto: largest active set that takes the time-parallel path (analytic-heavy modes)
  double* d_dbg = nullptr;  // debug trace [dbg_records][kDebugFields][ld], allocated by the first solve with params.debug
  int dbg_records = 0;
  bool dbg_valid = false;
  int ensure_debug_trace(int records);
  int concurrency_hint = 1;  // independent solves expected in flight on this device: the lane mappings share the device with them
  int sweep_lanes_max = 16384;  // largest active set whose Riccati sweep runs with the lanes of a problem sharing a step
  int sweep_wide_max = 0;       // ... with one lane per matrix entry (RiccatiWide): off, measured no faster (MAS_B200_SWEEP_WIDE=1 enables)
  int ensure_deriv_store(int block_doubles);
  bool coop_store = false;  // mas_b200_batch_set_trial_store(b, 2): trial store in the cooperative kernel too
  int ensure_trial_store(long long min_slots);
  int begin_download(double* X, double* U, double* cost, int* iterations, int* status);
  int wait_download();
  // result sink (mas_b200_batch_set_result_sink): page-locked host buffers that every solve streams its results into
  // while it runs -- after every iteration a kernel on `export_stream` writes the rows of the problems that have just
  // left the active set straight to the host (device-mapped addresses below), concurrently with the next iterations
  double *sink_X = nullptr, *sink_U = nullptr, *sink_cost = nullptr;
  int *sink_iters = nullptr, *sink_status = nullptr;
  bool sink_set = false, export_pending = false;
  cudaStream_t export_stream = nullptr;
  cudaEvent_t ev_exported = nullptr, ev_export_src = nullptr;
  int set_result_sink(double* X, double* U, double* cost, int* iterations, int* status);
  // mode 0: problems whose iteration counter equals `finished_at` and that are final; 1: status TIME_LIMIT; 2: all
  int export_results(int mode, int finished_at, int max_iterations, cudaEvent_t after);
  int finish_exports();   // end of a solve: event for wait_download() and for the next writer of X / U
  int fence_exports();    // the context stream waits for the last solve's exports before X / U are overwritten
  // strategy scratch (allocated on demand)
  double *d_U_old = nullptr, *d_X_old = nullptr, *d_cost_old = nullptr, *d_radius = nullptr;
  int* d_accepted = nullptr;
  int* h_counts = nullptr;      // pinned + mapped, [4]: active counts published by publish_count_kernel
  int* h_counts_dev = nullptr;  // the device-side address of the same words
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool per_problem_params = false;
  int tune_L = 0, tune_C = 0;
  int ls_mode = 0;  // 0 auto, 1 concurrent lanes (forward_kernel), 2 compacted rounds (trial_round_kernel + finish_kernel)
  // compacted-rounds line search scratch
  int* d_ls_list[2] = {nullptr, nullptr};
  int* d_round_count = nullptr;  // [8]
  int* d_accept_idx = nullptr;   // [ld], -1 = none
  double* d_accept_merit = nullptr;
  int rounds_used = 0;
  // optional per-launch timing (CUDA events on the launching stream), see mas_b200_batch_set_profiling
  bool profiling = false;
  struct TimedLaunch {
    cudaEvent_t e0, e1;
    int kind;  // 0 prologue, 1 backward, 2 forward
  };
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
  mas_b200_profile profile{};
  int* d_count_hist = nullptr;  // active problems at the start of every iteration of the last solve
  int hist_capacity = 0;
  void prof_begin(int kind);
  void prof_end();
  int prof_collect(int trips);
  mas_b200_batch_stats stats{};
  int last_L = 0, last_C = 0;

  virtual ~BatchBase();
  int allocate();
  int upload_rows(const double* host, double* dev, int rows);      // [batch][rows] host -> [rows][ld]
  int download_rows(const double* dev, double* host, int rows);    // [rows][ld] -> [batch][rows] host
  int ensure_strategy_scratch();
  // augmented-Lagrangian solver state of constrained models: multipliers [T][NC][ld], penalty [ld]; `al_fresh` =
  // next solve starts like a newly constructed solver after set_params (multipliers 0, penalty from the params)
  double *d_lam_eq = nullptr, *d_lam_ineq = nullptr, *d_penalty = nullptr;
  int neq = 0, nineq = 0;
  bool al_fresh = true;
  int allocate_constraint_state();
  int prepare_constraint_state(const mas_b200_ilqr_params& prm);
  virtual int initialize() = 0;
  virtual int solve(const mas_b200_ilqr_params& prm) = 0;
  virtual int rollout_all() = 0;                 // X, cost from U
  virtual int trust_region_step() = 0;           // uses d_*_old, d_radius, d_accepted
  // LineSearchNashStrategy pieces; d_U_cand, d_base_cost, d_ls_state allocated by ensure_strategy_scratch
  double *d_U_cand = nullptr, *d_base_cost = nullptr;
  int* d_ls_state = nullptr;
  int nash_ls_reduce(int n_scenarios, int n_agents, int phase);
  virtual int nash_ls_trial(int n_agents, double alpha) = 0;
  virtual int nash_ls_restore(int n_agents) = 0;
  int collect_stats();
};

inline int div_up(int a, int b) { return (a + b - 1) / b; }

template <class M>
struct BatchImpl : BatchBase {
  using View = BatchView<M::NX, M::NU>;
  View view{};
  BatchImpl() {
    neq = M::NEQ;
    nineq = M::NINEQ;
  }

  void make_view() {
    view.ld = ld;
    view.T = T;
    view.set_dt(desc.dt);
    view.deriv_mask = desc.deriv_mask;
    view.set_bounds(desc.has_input_bounds, desc.input_lower, desc.input_upper);
    view.per_problem_params = per_problem_params ? 1 : 0;
    for (int i = 0; i < kMaxParams; ++i) view.shared_p[i] = i < desc.num_params ? desc.params[i] : 0.0;
    view.params = d_params;
    view.x0 = d_x0;
    view.X = d_X;
    view.U = d_U;
    view.K = d_K;
    view.kff = d_k;
    view.cost = d_cost;
    view.merit = d_merit;
    view.iters = d_iters;
    view.status = d_status;
    view.trials = d_trials;
    view.reg_retries = d_reg;
    view.tolerance = 0.0;
    view.max_iterations = 0;
    view.lam_eq = d_lam_eq;
    view.lam_ineq = d_lam_ineq;
    view.penalty = d_penalty;
    view.penalty_increase = 5.0;
    view.constraint_tolerance = 1e-4;
    view.activation_tolerance = 1e-6;
    view.trial_X = trial_store ? d_trial_X : nullptr;
    view.trial_U = trial_store ? d_trial_U : nullptr;
    view.trial_slots = trial_store ? trial_slots : 0;
    view.dbg = dbg_valid ? d_dbg : nullptr;
    view.dbg_records = dbg_valid ? dbg_records : 0;
  }

  void apply_al_params(const mas_b200_ilqr_params& prm) {
    view.penalty_increase = prm.penalty_increase;
    view.constraint_tolerance = prm.constraint_tolerance;
    view.activation_tolerance = prm.inequality_activation_tolerance;
  }

  int launch_prologue(int max_iterations, const mas_b200_ilqr_params* prm = nullptr) {
    make_view();
    if (prm) apply_al_params(*prm);
    prologue_kernel<M><<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(view, batch, d_list[0], d_count, max_iterations);
    stats.kernel_launches++;
    MAS_CUDA_CHECK(cudaGetLastError());
    return MAS_B200_OK;
  }

  int initialize() override { return launch_prologue(0); }

  int rollout_all() override {
    make_view();
    rollout_kernel<M><<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(view, batch);
    stats.kernel_launches++;
    MAS_CUDA_CHECK(cudaGetLastError());
    return MAS_B200_OK;
  }

  int trust_region_step() override {
    make_view();
    trust_region_kernel<M><<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(view, batch, d_U_old, d_X_old, d_cost_old, d_radius, d_accepted);
    stats.kernel_launches++;
    MAS_CUDA_CHECK(cudaGetLastError());
    return MAS_B200_OK;
  }

  int nash_ls_trial(int n_agents, double alpha) override {
    make_view();
    nash_ls_trial_kernel<M><<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(view, batch, n_agents, d_U_old, d_U_cand, d_ls_state, alpha);
    stats.kernel_launches++;
    MAS_CUDA_CHECK(cudaGetLastError());
    return MAS_B200_OK;
  }

  int nash_ls_restore(int n_agents) override {
    make_view();
    nash_ls_restore_kernel<M><<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(view, batch, n_agents, d_U_old, d_X_old, d_cost_old, d_ls_state);
    stats.kernel_launches++;
    MAS_CUDA_CHECK(cudaGetLastError());
    return MAS_B200_OK;
  }

  template <int MASK_CT>
  void launch_time_parallel(int n_upper, int cur, int G) {
    const int n_pad = div_up(n_upper, 32) * 32;
    const long long threads = static_cast<long long>(n_pad) * G * (T + 1);
    linearize_kernel<M, MASK_CT><<<static_cast<int>((threads + kLinBlock - 1) / kLinBlock), kLinBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur,
                                                                                                                        d_deriv, deriv_cap, n_pad, G);
    // the recursion itself: lanes of a problem share a step while that still leaves most of the device idle (latency
    // bound), one thread per problem otherwise; models with path constraints always take the one-thread step
    using RL = RiccatiLanes<M, MASK_CT>;
    static const int env_sweep = std::getenv("MAS_B200_SWEEP_LANES") ? std::atoi(std::getenv("MAS_B200_SWEEP_LANES")) : -1;
    const bool lanes_ok = !HasConstraints<M>::value && M::NX <= 8;
    const bool use_lanes = lanes_ok && (env_sweep >= 0 ? env_sweep != 0 : static_cast<long long>(n_upper) * concurrency_hint <= sweep_lanes_max);
    // one lane per matrix entry (RiccatiWide): opt-in; on B200 one problem 0.834 vs 0.836 ms, 256-8,192 problems 3-5 % slower
    // than the four-lane sweep (profiles/r02_wide_probe.jsonl) -- the barriers between its six phases cost what the shorter
    // chains save
    using RW = RiccatiWide<M, MASK_CT>;
    static const int env_wide = std::getenv("MAS_B200_SWEEP_WIDE") ? std::atoi(std::getenv("MAS_B200_SWEEP_WIDE")) : -1;
    const bool use_wide = RW::kSupported && (env_wide >= 0 ? env_wide != 0 : static_cast<long long>(n_upper) * concurrency_hint <= sweep_wide_max);
    if (use_wide) {
      launch_sweep_wide<MASK_CT>(n_upper, cur);
    } else if (use_lanes) {
      constexpr int PW = 32 / RL::LG;
      riccati_sweep_lanes_kernel<M, MASK_CT><<<div_up(n_upper, PW), 32, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1), d_deriv,
                                                                                           deriv_cap);
    } else {
      riccati_sweep_kernel<M, MASK_CT><<<div_up(n_upper, kSweepBlock), kSweepBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur,
                                                                                                     d_count + (cur ^ 1), d_deriv, deriv_cap);
    }
    stats.kernel_launches += 2;
  }

  template <int MASK_CT>
  void launch_sweep_wide(int n_upper, int cur) {
    using RW = RiccatiWide<M, MASK_CT>;
    if constexpr (RW::kSupported) {
      constexpr int PW = 32 / RW::LW;
      riccati_sweep_wide_kernel<M, MASK_CT><<<div_up(n_upper, PW), 32, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1), d_deriv,
                                                                                          deriv_cap);
    }
  }

  void launch_backward(int n_upper, int cur) {
    const unsigned mask = desc.deriv_mask;
    // MAS_B200_BACKWARD_MODE / MAS_B200_TP_MAX: A/B switches for the entry points that own their batch (strategies)
    static const int env_mode = std::getenv("MAS_B200_BACKWARD_MODE") ? std::atoi(std::getenv("MAS_B200_BACKWARD_MODE")) : 0;
    static const int env_tp_max = std::getenv("MAS_B200_TP_MAX") ? std::atoi(std::getenv("MAS_B200_TP_MAX")) : 0;
    const int backward_mode = this->backward_mode ? this->backward_mode : env_mode;
    const int tp_max_problems = env_tp_max > 0 ? env_tp_max : this->tp_max_problems;
    {
      const bool fd_heavy_tp = !(mask & D_LXX) && (mask == 0u || mask != M::EXAMPLE_MASK);
      const bool want = backward_mode == 3 || (backward_mode == 0 && ls_mode == 0 && tune_L == 0 && (fd_heavy_tp || static_cast<long long>(n_upper) * concurrency_hint <= tp_max_problems));
      if (want && ensure_deriv_store(DerivBlock<M>::size) == MAS_B200_OK && n_upper <= deriv_cap) {
        const int G = fd_heavy_tp ? 8 : 2;
        if (mask == M::EXAMPLE_MASK) launch_time_parallel<static_cast<int>(M::EXAMPLE_MASK)>(n_upper, cur, G);
        else if (mask == 0u) launch_time_parallel<0>(n_upper, cur, G);
        else launch_time_parallel<-1>(n_upper, cur, G);
        return;
      }
    }
    // Many finite-difference callbacks (at least the n x n stage Hessian) and enough tasks to keep eight lanes busy
    // (n = 4: 40; the pendulum's 13 and the rocket's 21 measured slower than one thread per problem), and few enough
    // problems for eight lanes each to be resident: deal the stencil points out to the lanes (backward_lanes_kernel).
    const bool fd_heavy = !(mask & D_LXX) && (mask == 0u || mask != M::EXAMPLE_MASK) && DerivBlock<M>::n_tasks >= 32;
    if (backward_mode != 1 && (backward_mode == 2 || (backward_lanes_enabled && fd_heavy && ls_mode == 0 && tune_L == 0))) {
      constexpr int LB = 8;
      const size_t sm = (kBlock / LB) * DerivBlock<M>::size * sizeof(double);
      if (!resident_backward_lanes) {
        int blocks = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, backward_lanes_kernel<M, 0, LB>, kBlock, sm) != cudaSuccess || blocks <= 0) blocks = 4;
        resident_backward_lanes = static_cast<long long>(blocks) * kBlock * ctx->sm_count;
      }
      if (backward_mode == 2 || static_cast<long long>(n_upper) * LB <= resident_backward_lanes) {
        const int lgrid = static_cast<int>((static_cast<long long>(n_upper) * LB + kBlock - 1) / kBlock);
        if (mask == 0u)
          backward_lanes_kernel<M, 0, LB><<<lgrid, kBlock, sm, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1));
        else
          backward_lanes_kernel<M, -1, LB><<<lgrid, kBlock, sm, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1));
        stats.kernel_launches++;
        return;
      }
    }
    const int grid = div_up(n_upper, kBlock);
    const size_t smem = static_cast<long long>(grid) * kBlock <= kStageMaxThreads ? 2 * (M::NX + M::NU) * kStageStride * sizeof(double) : 0;
    if (mask == M::EXAMPLE_MASK)
      backward_kernel<M, static_cast<int>(M::EXAMPLE_MASK)><<<grid, kBlock, smem, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1));
    else if (mask == 0u)
      backward_kernel<M, 0><<<grid, kBlock, smem, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1));
    else
      backward_kernel<M, -1><<<grid, kBlock, smem, ctx->stream>>>(view, d_list[cur], d_count + cur, d_count + (cur ^ 1));
    stats.kernel_launches++;
  }

  template <int L, int C>
  void launch_forward_lc(int n_upper, int cur) {
    const long long threads = static_cast<long long>(n_upper) * L;
    const int grid = static_cast<int>((threads + kBlock - 1) / kBlock);
    const bool stage = L >= 4 && static_cast<long long>(grid) * kBlock <= kStageMaxThreads;
    const size_t smem = stage ? 2 * StepOperands<M::NX, M::NU>::NV * kStageStride * sizeof(double) : 0;
    forward_kernel<M, L, C><<<grid, kBlock, smem, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1));
    stats.kernel_launches++;
  }

  void launch_forward(int n_upper, int cur, int L, int C) {
    if (L == 1 && C == 2) launch_forward_lc<1, 2>(n_upper, cur);
    else if (L == 1) launch_forward_lc<1, 1>(n_upper, cur);
    else if (L == 2 && C == 2) launch_forward_lc<2, 2>(n_upper, cur);
    else if (L == 2) launch_forward_lc<2, 1>(n_upper, cur);
    else if (L == 4 && C == 2) launch_forward_lc<4, 2>(n_upper, cur);
    else if (L == 4) launch_forward_lc<4, 1>(n_upper, cur);
    else if (L == 8 && C == 2) launch_forward_lc<8, 2>(n_upper, cur);
    else if (L == 8) launch_forward_lc<8, 1>(n_upper, cur);
    else launch_forward_lc<16, 1>(n_upper, cur);
  }

  // Line search as compacted rounds of two step sizes each, then commit + stop test.
  int launch_rounds(int n_upper, int cur) {
    constexpr int C = 2;
    constexpr int R = (kNumAlphas + C - 1) / C;
    const int grid = div_up(n_upper, kBlock);
    MAS_CUDA_CHECK(cudaMemsetAsync(d_round_count, 0, 8 * sizeof(int), ctx->stream));
    for (int r = 0; r < R; ++r) {
      const int* in_list = r == 0 ? d_list[cur] : d_ls_list[(r - 1) & 1];
      const int* in_count = r == 0 ? d_count + cur : d_round_count + r;
      trial_round_kernel<M, C><<<grid, kBlock, 0, ctx->stream>>>(view, in_list, in_count, d_ls_list[r & 1], d_round_count + r + 1, r * C,
                                                                d_accept_idx, d_accept_merit);
      stats.kernel_launches++;
    }
    finish_kernel<M><<<grid, kBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1), d_accept_idx,
                                                      d_accept_merit);
    stats.kernel_launches++;
    rounds_used++;
    return MAS_B200_OK;
  }

  long long resident_backward_lanes = 0;  // resident threads of backward_lanes_kernel, queried at first use
  // Resident lanes of every forward_kernel variant on this device (occupancy x SMs), queried once.
  long long resident_lanes[5] = {0, 0, 0, 0, 0};  // L = 1 (C=2), 2, 4, 8, 16
  template <int L, int C>
  long long query_resident() const {
    int blocks = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, forward_kernel<M, L, C>, kBlock, 0) != cudaSuccess || blocks <= 0) blocks = 4;
    return static_cast<long long>(blocks) * kBlock * ctx->sm_count;
  }
  void query_occupancy() {
    if (resident_lanes[0]) return;

    resident_lanes[0] = query_resident<1, 2>();
    resident_lanes[1] = query_resident<2, 1>();
    resident_lanes[2] = query_resident<4, 2>();
    resident_lanes[3] = query_resident<8, 2>();
    resident_lanes[4] = query_resident<16, 1>();
  }

  // Lanes per problem: the widest mapping whose lanes all fit on the device at once.  Every kernel
  // here is bound by the latency of the T sequential time steps, so a launch that needs a second
  // wave costs a whole extra pass; within one wave more lanes per problem mean fewer step sizes per
  // lane and a shorter pass.  When even one lane per problem overflows the device, one lane it is.
  void choose_forward(int n_active, int* L, int* C) {
    int l = tune_L, c = tune_C;
    if (l == 0) {
      query_occupancy();
      l = 1;
      for (int k = 4; k >= 1; --k)
        if (static_cast<long long>(n_active) * (1 << k) * concurrency_hint <= resident_lanes[k]) {
          l = 1 << k;
          break;
        }
    }
    // two step sizes per lane where that saves passes: 4 lanes need 2 passes instead of 3 (0.48 vs 0.54 ms on the
    // launch list of the headline batch), 8 lanes 1 instead of 2 (0.25 vs 0.35 ms); with 2 lanes 3 two-chain passes
    // cost as much as 5 single ones (0.89 vs 0.86 ms), so they stay single
    if (c == 0) c = (l == 4 || l == 8 || l == 1) ? 2 : 1;
    if (l == 16) c = 1;
    *L = l;
    *C = c;
  }

  int solve(const mas_b200_ilqr_params& prm) override {
    using clock = std::chrono::steady_clock;
    const auto start = clock::now();
    int rc = fence_exports();
    if (rc) return rc;
    rc = prepare_constraint_state(prm);
    if (rc) return rc;
    dbg_valid = false;
    if (prm.debug) {
      rc = ensure_debug_trace(prm.max_iterations + 1);
      if (rc) return rc;
    }
    if (trial_store && prm.max_iterations > 0) {
      query_occupancy();
      // the widest launches that use the store: forward_kernel<L >= 4> with all lanes resident, or 16 lanes per problem
      long long need = 0;
      for (int k = 2; k <= 4; ++k) need = std::max(need, (k == 4 ? 1 : 2) * (resident_lanes[k] + kBlock));  // two chains for L = 4, 8
      need = std::min(need, 16ll * div_up(batch, kBlock) * kBlock);
      if (coop_store) need = std::max(need, 2ll * div_up(batch, kBlock) * kBlock);
      ensure_trial_store(need);
    }
    prof_begin(0);
    rc = launch_prologue(prm.max_iterations, &prm);
    prof_end();
    if (rc) return rc;
    if (profiling && prm.max_iterations > hist_capacity) {
      if (d_count_hist) cudaFree(d_count_hist);
      MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_count_hist), prm.max_iterations * sizeof(int)));
      hist_capacity = prm.max_iterations;
    }
    view.tolerance = prm.tolerance;
    view.max_iterations = prm.max_iterations;
    apply_al_params(prm);
    query_occupancy();
    const bool timed = std::isfinite(prm.max_ms);
    int n_upper = prm.max_iterations > 0 ? batch : 0;
    int cur = 0;
    int trips = 0;
    for (int it = 0; it < prm.max_iterations && n_upper > 0; ++it) {
      if (timed) {
        // the reference checks an integer-millisecond clock at the top of every iteration
        // (ilqr.hpp:84-90); here the budget covers the whole batch, so drain the stream first
        MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        const double elapsed_ms =
            static_cast<double>(std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - start).count());
        if (elapsed_ms > prm.max_ms) {
          rc = mark_time_limit(cur);
          if (rc) return rc;
          if (sink_set && (rc = export_results(1, 0, prm.max_iterations, nullptr))) return rc;
          break;
        }
      }
      int L, C;
      choose_forward(n_upper, &L, &C);
      last_L = L;
      last_C = C;
      prof_begin(1);
      launch_backward(n_upper, cur);
      prof_end();
      prof_begin(2);
      // Large active sets: compacted rounds (full warps, work stops at the first improving step size).
      // Small ones, where a pass is pure latency: all step sizes at once on L lanes per problem.
      // (measured on B200, 65,536 ST-lane problems: every round pays the latency of T sequential steps,
      //  so rounds lost to the lane mapping by 57 % per solve when both were measured; auto therefore = lanes)
      const bool rounds = ls_mode == 2;
      if (rounds) {
        rc = launch_rounds(n_upper, cur);
        if (rc) return rc;
        last_L = 0;
        last_C = 2;
      } else if (ls_mode == 3 || (ls_mode == 0 && tune_L == 0 && L == 1)) {
        // more problems than the device holds lanes: warp-cooperative search (32 problems per warp)
        // two step sizes per lane by default: +3% with several solves in flight, neutral alone (B200, 65,536 problems)
        const int cgrid = div_up(n_upper, kBlock);
        const bool cstore = coop_store && view.trial_X != nullptr && 2ll * cgrid * kBlock <= view.trial_slots;
        const bool sp = !view.per_problem_params;
        if (tune_C != 1 && cstore)
          forward_coop_kernel<M, 2, true, false><<<cgrid, kBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1));
        else if (tune_C != 1 && sp)
          forward_coop_kernel<M, 2, false, true><<<cgrid, kBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1));
        else if (tune_C != 1)
          forward_coop_kernel<M, 2, false, false><<<cgrid, kBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1));
        else
          forward_coop_kernel<M, 1, false, false><<<cgrid, kBlock, 0, ctx->stream>>>(view, d_list[cur], d_count + cur, d_list[cur ^ 1], d_count + (cur ^ 1));
        stats.kernel_launches++;
        last_L = 32;
        last_C = tune_C != 1 ? 2 : 1;
      } else {
        launch_forward(n_upper, cur, L, C);
      }
      prof_end();
      MAS_CUDA_CHECK(cudaGetLastError());
      if (profiling)
        MAS_CUDA_CHECK(cudaMemcpyAsync(d_count_hist + it, d_count + cur, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
      // the new active count goes to the host as a store into mapped pinned memory, not as a 4-byte D2H copy: the
      // copy engine is a FIFO shared with other contexts' result downloads, and a count queued behind a 170 MB
      // transfer stalls this loop for milliseconds (measured: -23 % solve throughput of four contexts next to one D2H stream)
      // (MAS_B200_COUNT_VIA_MEMCPY=1 restores the copy, for tools/copy_interference2.py)
      static const bool count_via_memcpy = std::getenv("MAS_B200_COUNT_VIA_MEMCPY") != nullptr;
      if (count_via_memcpy) {
        MAS_CUDA_CHECK(cudaMemcpyAsync(h_counts + (it & 1), d_count + (cur ^ 1), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      } else {
        publish_count_kernel<<<1, 1, 0, ctx->stream>>>(d_count + (cur ^ 1), h_counts_dev + (it & 1));
        stats.kernel_launches++;
      }
      MAS_CUDA_CHECK(cudaEventRecord(ev[it & 1], ctx->stream));
      // the problems that left the active set in this iteration travel to the host while the next iterations run
      if (sink_set && (rc = export_results(0, it + 1, prm.max_iterations, ev[it & 1]))) return rc;
      if (it >= 1) {
        // active count after iteration it-1: an upper bound for iteration it+1 (the list only shrinks)
        MAS_CUDA_CHECK(cudaEventSynchronize(ev[(it - 1) & 1]));
        n_upper = h_counts[(it - 1) & 1];
      }
      cur ^= 1;
      ++trips;
    }
    if (sink_set) {
      if (prm.max_iterations <= 0 && (rc = export_results(2, 0, 0, nullptr))) return rc;  // the prologue's rollout is the result
      if ((rc = finish_exports())) return rc;
    }
    stats.outer_iterations_run = trips;
    stats.forward_lanes = last_L;
    stats.forward_chains = last_C;
    if (profiling) return prof_collect(trips);
    return MAS_B200_OK;
  }

  int mark_time_limit(int cur);
};

__global__ void time_limit_kernel(const int* __restrict__ list, const int* __restrict__ count, int* status);

template <class M>
int BatchImpl<M>::mark_time_limit(int cur) {
  time_limit_kernel<<<div_up(batch, kBlock), kBlock, 0, ctx->stream>>>(d_list[cur], d_count + cur, d_status);
  stats.kernel_launches++;
  MAS_CUDA_CHECK(cudaGetLastError());
  return MAS_B200_OK;
}

int centralized_mixed_entry(Context* ctx, int n_blocks, const int* model_ids, int T, double dt, int has_bounds, const double* lo, const double* hi,
                            const mas_b200_ilqr_params& prm, int S, const double* x0, const double* params, double* X, double* U, double* costs, int* ints,
                            long long* launches);
int mixed_global_eval(Context* ctx, const int* model_ids, const int* state_offsets, const int* control_offsets, const double* params, int n_blocks,
                      int total_x, int total_u, const double* X, const double* U, int t, double* dyn_out, double* stage_out, double* terminal_out);

// factories, one translation unit per model (model_*.cu)
BatchBase* make_batch_st_lane();
BatchBase* make_batch_st_circ();
BatchBase* make_batch_lqr4();
BatchBase* make_batch_pendulum();
BatchBase* make_batch_rocket();
BatchBase* make_batch_st_lane_con();

}  // namespace mas_b200
