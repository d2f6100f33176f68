// centralized_host.cuh -- kernel wrapper and host launcher of the stacked (centralized) solve.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "centralized.cuh"
#include "engine.cuh"

namespace mas_b200 {

#ifndef MAS_CENTRALIZED_THREADS
#define MAS_CENTRALIZED_THREADS 512  /* 256 -> 512: 23.5 -> 18.5 ms per 32-agent scenario (4 warps per scheduler hide the latency of the barrier-separated phases); 640 / 768 / 1024: 19.9 / 21.2 / 23.5 ms */
#endif
constexpr int kCentralizedThreads = MAS_CENTRALIZED_THREADS;

// Persistent CTAs (one per SM slot the launch gets) pull scenarios from an atomic queue: `base` holds the pointers of
// scenario 0 / CTA 0; inputs and results (x0, prm, X, U, out_*) are indexed by scenario, the scratch (trial trajectories,
// gains, workspace) by CTA -- a run of thousands of scenarios needs the scratch of the resident CTAs only.  Scenarios differ
// widely in iteration count (4 to 88 with jittered track radii), so a CTA that finishes early takes the next one.
template <class M>
__global__ void __launch_bounds__(kCentralizedThreads) centralized_kernel(StackedProblem<M> base, int n_scenarios, size_t work_stride,
                                                                          size_t fast_offset, int fast_in_shared, int* queue) {
  extern __shared__ double mas_fast_scratch[];
  __shared__ int s_next;
  constexpr int NPs = (M::NP > 0 ? M::NP : 1);
  const int ns = base.A * M::NX, ms = base.A * M::NU, T = base.T;
  const int cta = blockIdx.x;
  for (;;) {
    __syncthreads();  // everybody is done with the previous scenario (and with s_next)
    if (threadIdx.x == 0) s_next = atomicAdd(queue, 1);
    __syncthreads();
    const int s = s_next;
    if (s >= n_scenarios) return;
    StackedProblem<M> P = base;
    P.x0 = base.x0 + static_cast<size_t>(s) * ns;
    P.prm = base.prm + static_cast<size_t>(s) * base.A * NPs;
    P.X = base.X + static_cast<size_t>(s) * (T + 1) * ns;
    P.U = base.U + static_cast<size_t>(s) * T * ms;
    P.out_cost = base.out_cost + static_cast<size_t>(s) * (1 + base.A);
    P.out_int = base.out_int + static_cast<size_t>(s) * 4;
    P.Xt = base.Xt + static_cast<size_t>(cta) * (T + 1) * ns;
    P.Ut = base.Ut + static_cast<size_t>(cta) * T * ms;
    P.K = base.K + static_cast<size_t>(cta) * T * ms * ns;
    P.kff = base.kff + static_cast<size_t>(cta) * T * ms;
    P.work = base.work + static_cast<size_t>(cta) * work_stride;
    P.fast = mas_fast_scratch;
    if (s != 0) P.phase_cycles = nullptr;  // diagnostics: scenario 0 only
    if (fast_in_shared) {
      stacked_solve<M, true>(P, threadIdx.x, blockDim.x);
    } else {  // stacked problems too large for shared memory: the same scratch inside the global workspace
      P.fast = P.work + fast_offset;
      stacked_solve<M, false>(P, threadIdx.x, blockDim.x);
    }
  }
}

// CentralizedStrategy::operator() for n_scenarios scenarios of n_agents agents (host arrays as in
// mas_b200_strategy_run).  trace_iterations (optional): [scenario] iterations of the stacked solve.
template <class M>
int run_centralized(Context* ctx, const mas_b200_ocp_desc& d, const mas_b200_ilqr_params& prm, int S, int A, const double* x0,
                    const double* model_params, const double* U_init, double* X, double* U, double* costs, double* total_cost,
                    int* iterations_out, int* status_out, long long* launches) {
  constexpr int NX = M::NX, NU = M::NU, NPs = (M::NP > 0 ? M::NP : 1);
  const int T = d.horizon_steps, ns = A * NX, ms = A * NU;
  const StackedWork W(A, NX, NU);
  cudaStream_t st = ctx->stream;
  MAS_CUDA_CHECK(cudaSetDevice(ctx->device));
  // host staging in the stacked layout: X [S][T+1][ns], U [S][T][ms], params [S][A][NPs]
  std::vector<double> hU(static_cast<size_t>(S) * T * ms, 0.0), hP(static_cast<size_t>(S) * A * NPs, 0.0);
  for (int s = 0; s < S; ++s)
    for (int a = 0; a < A; ++a) {
      for (int i = 0; i < M::NP; ++i)
        hP[(static_cast<size_t>(s) * A + a) * NPs + i] = model_params ? model_params[(static_cast<size_t>(s) * A + a) * M::NP + i] : d.params[i];
    }
  // U_init is deliberately not used: build_global_ocp() never sets initial_controls, so initialize_problem() starts the
  // stacked OCP from zero controls whatever the agents hold (multi_agent_problem.hpp:52-127, ocp.hpp:104-108)
  (void)U_init;
  struct Buffers {
    double *x0 = nullptr, *prm = nullptr, *X = nullptr, *Xt = nullptr, *U = nullptr, *Ut = nullptr, *K = nullptr, *k = nullptr, *work = nullptr,
           *oc = nullptr;
    int *oi = nullptr, *queue = nullptr;
    ~Buffers() {
      for (double* p : {x0, prm, X, Xt, U, Ut, K, k, work, oc})
        if (p) cudaFree(p);
      if (oi) cudaFree(oi);
      if (queue) cudaFree(queue);
    }
  };
  // persistent CTAs: as many as the device holds at once (one per SM with the 200 KB shared-memory scratch), at most S
  const size_t fast_bytes_q = W.fast_doubles * sizeof(double);
  int max_optin_q = 0, per_sm = 1;
  MAS_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin_q, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  const bool shared_q = fast_bytes_q <= static_cast<size_t>(max_optin_q);
  if (shared_q) MAS_CUDA_CHECK(cudaFuncSetAttribute(centralized_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fast_bytes_q)));
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, centralized_kernel<M>, kCentralizedThreads, shared_q ? fast_bytes_q : 0) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const int G = std::min(S, ctx->sm_count * per_sm);
  const long long key[4] = {S, A, T, M::ID};
  const bool reuse = ctx->centralized_workspace && std::equal(key, key + 4, ctx->centralized_key);
  if (!reuse) {
    ctx->centralized_workspace.reset();  // free the old shape first
    ctx->centralized_workspace = std::make_shared<Buffers>();
    std::fill(ctx->centralized_key, ctx->centralized_key + 4, 0LL);  // valid only once every allocation succeeded
  }
  Buffers& b = *static_cast<Buffers*>(ctx->centralized_workspace.get());
  auto dalloc = [&](double** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(double)); };
  const size_t Ss = static_cast<size_t>(S);
  if (!reuse) {
  MAS_CUDA_CHECK(dalloc(&b.x0, Ss * ns));
  MAS_CUDA_CHECK(dalloc(&b.prm, Ss * A * NPs));
  MAS_CUDA_CHECK(dalloc(&b.X, Ss * (T + 1) * ns));
  const size_t Gs = static_cast<size_t>(G);
  MAS_CUDA_CHECK(dalloc(&b.Xt, Gs * (T + 1) * ns));
  MAS_CUDA_CHECK(dalloc(&b.U, Ss * T * ms));
  MAS_CUDA_CHECK(dalloc(&b.Ut, Gs * T * ms));
  MAS_CUDA_CHECK(dalloc(&b.K, Gs * T * ms * ns));
  MAS_CUDA_CHECK(dalloc(&b.k, Gs * T * ms));
  MAS_CUDA_CHECK(dalloc(&b.work, Gs * W.total));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&b.queue), sizeof(int)));
  MAS_CUDA_CHECK(dalloc(&b.oc, Ss * (1 + A)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&b.oi), Ss * 4 * sizeof(int)));
  std::copy(key, key + 4, ctx->centralized_key);
  }
  MAS_CUDA_CHECK(cudaMemcpyAsync(b.x0, x0, Ss * ns * sizeof(double), cudaMemcpyHostToDevice, st));  // [S][A][NX] is already [S][ns]
  MAS_CUDA_CHECK(cudaMemcpyAsync(b.prm, hP.data(), hP.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(b.U, hU.data(), hU.size() * sizeof(double), cudaMemcpyHostToDevice, st));

  StackedProblem<M> P{};
  P.A = A;
  P.T = T;
  P.dt = d.dt;
  P.has_bounds = d.has_input_bounds;
  for (int i = 0; i < NU; ++i) {
    P.lo[i] = d.input_lower[i];
    P.hi[i] = d.input_upper[i];
  }
  P.tolerance = prm.tolerance;
  P.max_iterations = prm.max_iterations;
  P.max_ms = prm.max_ms;  // per stacked solve, as for the reference's solve() of the global OCP
  P.x0 = b.x0;
  P.prm = b.prm;
  P.X = b.X;
  P.U = b.U;
  P.Xt = b.Xt;
  P.Ut = b.Ut;
  P.K = b.K;
  P.kff = b.k;
  P.work = b.work;
  P.out_cost = b.oc;
  P.out_int = b.oi;
  // opt-in, never the default, excluded from the parity gate: dense gain / value-update products on the fp64 tensor cores
  P.use_dmma = std::getenv("MAS_B200_CENTRALIZED_DMMA") && std::atoi(std::getenv("MAS_B200_CENTRALIZED_DMMA")) != 0;
  // MAS_B200_CENTRALIZED_PHASES=1: SM cycles per phase of scenario 0 to stderr (how the time of a stacked solve splits
  // between finite differences, the dense Riccati algebra and the line-search rollouts)
  long long* d_phase = nullptr;
  if (std::getenv("MAS_B200_CENTRALIZED_PHASES")) {
    MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_phase), kNumPhases * sizeof(long long)));
    MAS_CUDA_CHECK(cudaMemsetAsync(d_phase, 0, kNumPhases * sizeof(long long), st));
  }
  P.phase_cycles = d_phase;
  // K, Q_ux and the factorisation scratch live in shared memory when they fit (A = 32 single-track agents: 199 KB)
  const size_t fast_bytes = W.fast_doubles * sizeof(double);
  int max_optin = 0;
  MAS_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  const int in_shared = fast_bytes <= static_cast<size_t>(max_optin) ? 1 : 0;
  if (in_shared)
    MAS_CUDA_CHECK(cudaFuncSetAttribute(centralized_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fast_bytes)));
  MAS_CUDA_CHECK(cudaMemsetAsync(b.queue, 0, sizeof(int), st));
  centralized_kernel<M><<<G, kCentralizedThreads, in_shared ? fast_bytes : 0, st>>>(P, S, W.total, W.fast, in_shared, b.queue);
  if (launches) (*launches)++;
  MAS_CUDA_CHECK(cudaGetLastError());

  std::vector<double> hX(Ss * (T + 1) * ns), hUo(Ss * T * ms), hc(Ss * (1 + A));
  std::vector<int> hi(Ss * 4);
  MAS_CUDA_CHECK(cudaMemcpyAsync(hX.data(), b.X, hX.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(hUo.data(), b.U, hUo.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(hc.data(), b.oc, hc.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(hi.data(), b.oi, hi.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaStreamSynchronize(st));
  if (d_phase) {
    long long h[kNumPhases];
    MAS_CUDA_CHECK(cudaMemcpy(h, d_phase, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(d_phase);
    static const char* names[kNumPhases] = {"terminal FD", "stage FD tables + derivatives (not overlapped)", "Q assembly (A^T V A ...)", "LLT + inverse || FD of the next step", "(unused)",
                                            "gains", "value update", "rollouts (prologue + line search)"};
    long long tot = 0;
    for (long long c : h) tot += c;
    std::fprintf(stderr, "[mas_b200] centralized scenario 0: %lld cycles\n", tot);
    for (int i = 0; i < kNumPhases; ++i) std::fprintf(stderr, "[mas_b200]   %-36s %12lld  %5.1f %%\n", names[i], h[i], 100.0 * h[i] / (tot ? tot : 1));
  }
  // scatter the stacked trajectories back to the agents' blocks (centralized.hpp:30-31)
  for (int s = 0; s < S; ++s) {
    for (int a = 0; a < A; ++a) {
      const size_t ag = static_cast<size_t>(s) * A + a;
      if (X)
        for (int t = 0; t <= T; ++t)
          for (int i = 0; i < NX; ++i) X[(ag * (T + 1) + t) * NX + i] = hX[(static_cast<size_t>(s) * (T + 1) + t) * ns + a * NX + i];
      if (U)
        for (int t = 0; t < T; ++t)
          for (int i = 0; i < NU; ++i) U[(ag * T + t) * NU + i] = hUo[(static_cast<size_t>(s) * T + t) * ms + a * NU + i];
      if (costs) costs[ag] = hc[static_cast<size_t>(s) * (1 + A) + 1 + a];
    }
    if (total_cost) total_cost[s] = hc[static_cast<size_t>(s) * (1 + A)];
    if (iterations_out) iterations_out[s] = hi[static_cast<size_t>(s) * 4 + 0];
    if (status_out) status_out[s] = hi[static_cast<size_t>(s) * 4 + 1];
  }
  return MAS_B200_OK;
}

using CentralizedFn = int (*)(Context*, const mas_b200_ocp_desc&, const mas_b200_ilqr_params&, int, int, const double*, const double*, const double*,
                              double*, double*, double*, double*, int*, int*, long long*);
CentralizedFn centralized_entry(int model_id);

}  // namespace mas_b200
