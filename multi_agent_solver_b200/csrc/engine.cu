// engine.cu -- model-independent parts of the batch engine: HBM allocation, layout transposes,
// host<->device staging, statistics.
#include "engine.cuh"

namespace mas_b200 {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

// [batch][rows] (row-major, problem-major) -> [rows][ld]; 32x32 shared-memory tiles so both the
// global read (along rows) and the global write (along the batch) are coalesced.
__global__ void aos_to_soa_kernel(const double* __restrict__ src, double* __restrict__ dst, int batch, int rows, int ld) {
  __shared__ double tile[32][33];
  const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int b = b0 + j, r = r0 + threadIdx.x;
    if (b < batch && r < rows) tile[j][threadIdx.x] = src[static_cast<size_t>(b) * rows + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, b = b0 + threadIdx.x;
    if (b < batch && r < rows) dst[static_cast<size_t>(r) * ld + b] = tile[threadIdx.x][j];
  }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ src, double* __restrict__ dst, int batch, int rows, int ld) {
  __shared__ double tile[32][33];
  const int b0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, b = b0 + threadIdx.x;
    if (b < batch && r < rows) tile[j][threadIdx.x] = src[static_cast<size_t>(r) * ld + b];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int b = b0 + j, r = r0 + threadIdx.x;
    if (b < batch && r < rows) dst[static_cast<size_t>(b) * rows + r] = tile[threadIdx.x][j];
  }
}

__global__ void publish_count_kernel(const int* __restrict__ count, int* mapped_host_word) { *mapped_host_word = *count; }

__global__ void time_limit_kernel(const int* __restrict__ list, const int* __restrict__ count, int* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < *count) status[list[i]] = STATUS_TIME_LIMIT;
}

// Result sink: rows of finished problems straight to page-locked host memory (device-mapped), [batch][rows] row-major --
// per problem exactly the bytes of the reference's column-major best_states / best_controls.  A warp looks at 32
// neighbouring problems, then moves the finished ones one after the other: its lanes read 32 rows of the problem's
// [rows][ld] column and store them as one 256-byte run over PCIe.  mode 0: problems whose iteration counter equals
// `finished_at` and whose flag is final (they left the active set in the iteration that has just ended and are never
// written again); mode 1: problems stopped by the time budget; mode 2: all.
__global__ void export_results_kernel(const double* __restrict__ dX, const double* __restrict__ dU, const double* __restrict__ dcost,
                                      const int* __restrict__ iters, const int* __restrict__ status, int batch, int ld, int rowsX, int rowsU, int mode,
                                      int finished_at, int max_iterations, double* hX, double* hU, double* hcost, int* hiters, int* hstatus) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int groups = (batch + 31) >> 5;
  for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < groups; g += warps) {
    const int p = g * 32 + lane;
    bool done = false;
    if (p < batch) {
      const int it = iters[p], st = status[p];
      done = mode == 2 || (mode == 1 ? st == STATUS_TIME_LIMIT : (it == finished_at && (st == STATUS_CONVERGED || it >= max_iterations)));
      if (done) {
        if (hcost) hcost[p] = dcost[p];
        if (hiters) hiters[p] = it;
        if (hstatus) hstatus[p] = st;
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, done);
    while (todo) {
      const int q = g * 32 + (__ffs(todo) - 1);
      todo &= todo - 1;
      if (hX) {
        double* dst = hX + static_cast<size_t>(q) * rowsX;
        int r = lane;
        for (; r + 96 < rowsX; r += 128) {  // four independent loads in flight per lane
          const double a = __ldcs(dX + static_cast<size_t>(r) * ld + q), b = __ldcs(dX + static_cast<size_t>(r + 32) * ld + q),
                       c = __ldcs(dX + static_cast<size_t>(r + 64) * ld + q), d = __ldcs(dX + static_cast<size_t>(r + 96) * ld + q);
          dst[r] = a;
          dst[r + 32] = b;
          dst[r + 64] = c;
          dst[r + 96] = d;
        }
        for (; r < rowsX; r += 32) dst[r] = __ldcs(dX + static_cast<size_t>(r) * ld + q);
      }
      if (hU) {
        double* dst = hU + static_cast<size_t>(q) * rowsU;
        int r = lane;
        for (; r + 96 < rowsU; r += 128) {
          const double a = __ldcs(dU + static_cast<size_t>(r) * ld + q), b = __ldcs(dU + static_cast<size_t>(r + 32) * ld + q),
                       c = __ldcs(dU + static_cast<size_t>(r + 64) * ld + q), d = __ldcs(dU + static_cast<size_t>(r + 96) * ld + q);
          dst[r] = a;
          dst[r + 32] = b;
          dst[r + 64] = c;
          dst[r + 96] = d;
        }
        for (; r < rowsU; r += 32) dst[r] = __ldcs(dU + static_cast<size_t>(r) * ld + q);
      }
    }
  }
}

// 8 independent DFMA chains per thread; reports 2 flops per fma.
__global__ void dfma_probe_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c);
    a1 = fma(a1, m, c);
    a2 = fma(a2, m, c);
    a3 = fma(a3, m, c);
    a4 = fma(a4, m, c);
    a5 = fma(a5, m, c);
    a6 = fma(a6, m, c);
    a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// Brute-force check of the straight-line divisions of portable_math.h against the instructions they stand for, on
// operands drawn from a counter-based generator: out[0] pairs checked, out[1] pairs on which div_spec kept `exact`,
// out[2] of those whose bits differ from div.rn.f64, out[3] / out[4] the same two counts for div_const_spec against `/`
// (divisors 2.5 and 6).  Operand classes by (index mod 4): any 64-bit pattern; exponents within 2^+-40 of 1; the same
// with the numerator's low mantissa bits cleared (quotients near ties); numerator from a cos-like range.
__global__ void division_selftest_kernel(unsigned long long seed, int per_thread, unsigned long long* out) {
  auto mix = [](unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  const unsigned long long tid = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  unsigned long long n_exact = 0, n_bad = 0, nc_exact = 0, nc_bad = 0;
  for (int k = 0; k < per_thread; ++k) {
    const unsigned long long idx = tid * per_thread + k;
    unsigned long long ra = mix(seed + 2 * idx), rb = mix(seed + 2 * idx + 1);
    const int cls = static_cast<int>(idx & 3);
    if (cls != 0) {
      const unsigned long long ea = 1023 - 40 + (ra >> 52) % 81, eb = 1023 - 40 + (rb >> 52) % 81;
      ra = (ra & 0x800FFFFFFFFFFFFFull) | (ea << 52);
      rb = (rb & 0x800FFFFFFFFFFFFFull) | (eb << 52);
      if (cls == 2) ra &= ~0xFFFFFFFFull;
      if (cls == 3) rb = (rb & 0x800FFFFFFFFFFFFFull) | (1022ull << 52);
    }
    const double a = __longlong_as_double(static_cast<long long>(ra)), b = __longlong_as_double(static_cast<long long>(rb));
    double ref;
    asm volatile("div.rn.f64 %0, %1, %2;" : "=d"(ref) : "d"(a), "d"(b));
    bool exact = true;
    const double q = pm::div_spec(a, b, &exact);
    if (exact) {
      ++n_exact;
      if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) ++n_bad;
    }
    const double divisor = (idx & 4) ? 2.5 : 6.0, recip = (idx & 4) ? 1.0 / 2.5 : 1.0 / 6.0;
    double refc;
    asm volatile("div.rn.f64 %0, %1, %2;" : "=d"(refc) : "d"(a), "d"(divisor));
    bool exact_c = true;
    const double qc = pm::div_const_spec(a, divisor, recip, &exact_c);
    if (exact_c) {
      ++nc_exact;
      if (__double_as_longlong(qc) != __double_as_longlong(refc) && !(qc != qc && refc != refc)) ++nc_bad;
    }
  }
  atomicAdd(out + 0, static_cast<unsigned long long>(per_thread));
  atomicAdd(out + 1, n_exact);
  atomicAdd(out + 2, n_bad);
  atomicAdd(out + 3, nc_exact);
  atomicAdd(out + 4, nc_bad);
}

BatchBase::~BatchBase() {
  if (ctx) cudaSetDevice(ctx->device);
  // transfers that still read the buffers freed below
  if (download_pending) cudaEventSynchronize(ev_downloaded);
  if (export_pending) cudaEventSynchronize(ev_exported);
  double* dptrs[] = {d_x0, d_X, d_U, d_K, d_k, d_cost, d_merit, d_params, d_stage, d_U_old, d_X_old, d_cost_old, d_radius, d_accept_merit, d_U_cand, d_base_cost, d_lam_eq, d_lam_ineq, d_penalty};
  for (double* p : dptrs)
    if (p) cudaFree(p);
  int* iptrs[] = {d_iters, d_status, d_trials, d_reg, d_list[0], d_list[1], d_count, d_accepted, d_ls_list[0], d_ls_list[1], d_round_count, d_accept_idx, d_ls_state};
  for (int* p : iptrs)
    if (p) cudaFree(p);
  if (h_counts) cudaFreeHost(h_counts);
  if (d_count_hist) cudaFree(d_count_hist);
  if (export_stream) cudaStreamDestroy(export_stream);
  if (ev_exported) cudaEventDestroy(ev_exported);
  if (ev_export_src) cudaEventDestroy(ev_export_src);
  if (d_out_stage) cudaFree(d_out_stage);
  if (d_deriv) cudaFree(d_deriv);
  if (d_dbg) cudaFree(d_dbg);
  if (d_trial_X) cudaFree(d_trial_X);
  if (d_trial_U) cudaFree(d_trial_U);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (ev_staged) cudaEventDestroy(ev_staged);
  if (ev_downloaded) cudaEventDestroy(ev_downloaded);
  for (auto& tl : timed) {
    cudaEventDestroy(tl.e0);
    cudaEventDestroy(tl.e1);
  }
  for (auto& e : event_pool) cudaEventDestroy(e);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
}

int BatchBase::allocate() {
  MAS_CUDA_CHECK(cudaSetDevice(ctx->device));
  ld = ((batch + 31) / 32) * 32;
  const size_t L = static_cast<size_t>(ld);
  auto dalloc = [&](double** p, size_t n) -> cudaError_t {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, n * sizeof(double), ctx->stream);
    return e;
  };
  auto ialloc = [&](int** p, size_t n) -> cudaError_t {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(*p, 0, n * sizeof(int), ctx->stream);
    return e;
  };
  MAS_CUDA_CHECK(dalloc(&d_x0, L * nx));
  MAS_CUDA_CHECK(dalloc(&d_X, L * nx * (T + 1)));
  MAS_CUDA_CHECK(dalloc(&d_U, L * nu * T));
  MAS_CUDA_CHECK(dalloc(&d_K, L * nu * nx * T));
  MAS_CUDA_CHECK(dalloc(&d_k, L * nu * T));
  MAS_CUDA_CHECK(dalloc(&d_cost, L));
  MAS_CUDA_CHECK(dalloc(&d_merit, L));
  const size_t stage_rows = static_cast<size_t>(nx) * (T + 1) > static_cast<size_t>(nu) * T ? static_cast<size_t>(nx) * (T + 1) : static_cast<size_t>(nu) * T;
  MAS_CUDA_CHECK(dalloc(&d_stage, static_cast<size_t>(batch) * stage_rows));
  if (np > 0) MAS_CUDA_CHECK(dalloc(&d_params, L * np));
  MAS_CUDA_CHECK(ialloc(&d_iters, L));
  MAS_CUDA_CHECK(ialloc(&d_status, L));
  MAS_CUDA_CHECK(ialloc(&d_trials, L));
  MAS_CUDA_CHECK(ialloc(&d_reg, L));
  MAS_CUDA_CHECK(ialloc(&d_list[0], L));
  MAS_CUDA_CHECK(ialloc(&d_list[1], L));
  MAS_CUDA_CHECK(ialloc(&d_count, 2));
  MAS_CUDA_CHECK(ialloc(&d_ls_list[0], L));
  MAS_CUDA_CHECK(ialloc(&d_ls_list[1], L));
  MAS_CUDA_CHECK(ialloc(&d_round_count, 8));
  MAS_CUDA_CHECK(ialloc(&d_accept_idx, L));
  MAS_CUDA_CHECK(cudaMemsetAsync(d_accept_idx, 0xFF, L * sizeof(int), ctx->stream));  // -1: no accepted step size
  MAS_CUDA_CHECK(dalloc(&d_accept_merit, L));
  MAS_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&h_counts), 4 * sizeof(int), cudaHostAllocMapped));
  MAS_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_counts_dev), h_counts, 0));
  const unsigned ev_flags = cudaEventDisableTiming | (ctx->blocking_sync ? cudaEventBlockingSync : 0u);
  MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev[0], ev_flags));
  MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev[1], ev_flags));
  return allocate_constraint_state();
}

int BatchBase::allocate_constraint_state() {
  if (neq == 0 && nineq == 0) return MAS_B200_OK;
  if (T > kMaxALHorizon) {
    set_last_error("constrained models support horizons up to 128 steps");
    return MAS_B200_ERR_UNSUPPORTED;
  }
  const size_t L = static_cast<size_t>(ld);
  if (neq > 0) MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_lam_eq), L * neq * T * sizeof(double)));
  if (nineq > 0) MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_lam_ineq), L * nineq * T * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_penalty), L * sizeof(double)));
  // a default-constructed solver: zero multipliers, penalty 10 (ilqr.hpp:31); the first solve re-initialises both
  // from its params (`al_fresh`)
  mas_b200_ilqr_params defaults;
  mas_b200_ilqr_default_params(&defaults);
  al_fresh = true;
  const int rc = prepare_constraint_state(defaults);
  al_fresh = true;
  return rc;
}

// A fresh reference solver has zero multipliers and the penalty given to set_params (ilqr.hpp:31,47-48,331-338);
// afterwards both persist from solve to solve.
int BatchBase::prepare_constraint_state(const mas_b200_ilqr_params& prm) {
  if ((neq == 0 && nineq == 0) || !al_fresh) return MAS_B200_OK;
  const size_t L = static_cast<size_t>(ld);
  if (d_lam_eq) MAS_CUDA_CHECK(cudaMemsetAsync(d_lam_eq, 0, L * neq * T * sizeof(double), ctx->stream));
  if (d_lam_ineq) MAS_CUDA_CHECK(cudaMemsetAsync(d_lam_ineq, 0, L * nineq * T * sizeof(double), ctx->stream));
  std::vector<double> rho(L, prm.penalty);
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_penalty, rho.data(), L * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  al_fresh = false;
  return MAS_B200_OK;
}

// Derivative blocks of the time-parallel backward pass: [T+1][block_doubles][deriv_cap] doubles.  As many slots as the
// batch needs, within 512 MB (e.g. 13,600 problems at T = 80, the whole batch at T = 10); larger active sets take the
// fused kernels.  A failed allocation disables the path for this batch.
int BatchBase::ensure_deriv_store(int block_doubles) {
  if (d_deriv) return MAS_B200_OK;
  if (deriv_cap < 0) return MAS_B200_ERR_CUDA;
  const size_t per_slot = static_cast<size_t>(T + 1) * block_doubles * sizeof(double);
  long long cap = static_cast<long long>((512ull << 20) / per_slot);
  cap = std::min<long long>(cap, ld);
  cap = (cap / 32) * 32;
  if (cap < 32) cap = 32;
  if (cudaMalloc(reinterpret_cast<void**>(&d_deriv), per_slot * cap) != cudaSuccess) {
    cudaGetLastError();
    d_deriv = nullptr;
    deriv_cap = -1;
    return MAS_B200_ERR_CUDA;
  }
  deriv_cap = static_cast<int>(cap);
  return MAS_B200_OK;
}

// Trace buffer of `debug` solves; every record starts as NaN so that unused iterations are recognisable.
int BatchBase::ensure_debug_trace(int records) {
  if (records > dbg_records) {
    if (d_dbg) cudaFree(d_dbg);
    d_dbg = nullptr;
    dbg_records = 0;
    MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_dbg), static_cast<size_t>(records) * kDebugFields * ld * sizeof(double)));
    dbg_records = records;
  }
  MAS_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0xFF, static_cast<size_t>(dbg_records) * kDebugFields * ld * sizeof(double), ctx->stream));
  dbg_valid = true;
  return MAS_B200_OK;
}

int BatchBase::ensure_strategy_scratch() {
  if (d_U_old) return MAS_B200_OK;
  const size_t L = static_cast<size_t>(ld);
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_U_old), L * nu * T * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_X_old), L * nx * (T + 1) * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_cost_old), L * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_radius), L * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_accepted), L * sizeof(int)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_U_cand), L * nu * T * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_base_cost), L * sizeof(double)));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_ls_state), L * sizeof(int)));
  MAS_CUDA_CHECK(cudaMemsetAsync(d_ls_state, 0, L * sizeof(int), ctx->stream));
  return MAS_B200_OK;
}

// phase -1: base = joint cost (nash.hpp:103).  phase 0, after the Jacobi solve (:119-122,173-176): joint cost not
// lower than base -> state 1 (search), else base = joint cost, state 0.  phase 1, after a trial (:143-153):
// searching scenarios whose trial joint cost is below base accept it -> state 2.
__global__ void nash_ls_reduce_kernel(const double* __restrict__ cost, int n_scenarios, int n_agents, double* base_cost, int* state, int phase) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scenarios) return;
  if (phase == 1 && state[s] != 1) return;
  double c = 0.0;
  for (int a = 0; a < n_agents; ++a) c += cost[static_cast<size_t>(s) * n_agents + a];
  if (phase < 0) {
    base_cost[s] = c;
    state[s] = 0;
  } else if (phase == 0) {
    if (c >= base_cost[s]) {
      state[s] = 1;
    } else {
      base_cost[s] = c;
      state[s] = 0;
    }
  } else if (c < base_cost[s]) {
    base_cost[s] = c;
    state[s] = 2;
  }
}

// ---- stacked functions of a MultiAgentProblem whose agents are of different models (build_global_ocp,
// multi_agent_problem.hpp:94-125): block-diagonal dynamics, stage / terminal costs summed in block order from 0.0.
// One thread per block evaluates its agent's registered functors (run-time dispatch on the model id); thread 0 then
// forms the two sums in block order.
struct EvalBlock {
  int model_id, state_offset, control_offset;
  double params[kMaxParams];
};
template <class M>
__device__ void mixed_block_eval(const EvalBlock& b, const double* X, const double* U, int t, double* dyn, double* stage, double* terminal) {
  double x[M::NX], u[M::NU], d[M::NX];
  for (int i = 0; i < M::NX; ++i) x[i] = X[b.state_offset + i];
  for (int i = 0; i < M::NU; ++i) u[i] = U[b.control_offset + i];
  M::dynamics(x, u, b.params, d);
  for (int i = 0; i < M::NX; ++i) dyn[b.state_offset + i] = d[i];
  *stage = M::stage(x, u, t, b.params);
  *terminal = M::terminal(x, b.params);
}
__global__ void mixed_global_eval_kernel(const EvalBlock* blocks, int n_blocks, const double* X, const double* U, int t, double* dyn, double* terms,
                                         double* sums) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_blocks) {
    const EvalBlock b = blocks[a];
    double* st = terms + a;
    double* te = terms + n_blocks + a;
    switch (b.model_id) {
      case StLane::ID: mixed_block_eval<StLane>(b, X, U, t, dyn, st, te); break;
      case StCirc::ID: mixed_block_eval<StCirc>(b, X, U, t, dyn, st, te); break;
      case Lqr4::ID: mixed_block_eval<Lqr4>(b, X, U, t, dyn, st, te); break;
      case Pendulum::ID: mixed_block_eval<Pendulum>(b, X, U, t, dyn, st, te); break;
      case Rocket::ID: mixed_block_eval<Rocket>(b, X, U, t, dyn, st, te); break;
      case StLaneCon::ID: mixed_block_eval<StLaneCon>(b, X, U, t, dyn, st, te); break;
    }
  }
}
__global__ void mixed_global_sum_kernel(const double* terms, int n_blocks, double* sums) {
  double s = 0.0, e = 0.0;
  for (int a = 0; a < n_blocks; ++a) s += terms[a];
  for (int a = 0; a < n_blocks; ++a) e += terms[n_blocks + a];
  sums[0] = s;
  sums[1] = e;
}

int mixed_global_eval(Context* ctx, const int* model_ids, const int* state_offsets, const int* control_offsets, const double* params /* [n][kMaxParams] */,
                      int n_blocks, int total_x, int total_u, const double* X, const double* U, int t, double* dyn_out, double* stage_out,
                      double* terminal_out) {
  std::vector<EvalBlock> hb(n_blocks);
  for (int a = 0; a < n_blocks; ++a) {
    hb[a].model_id = model_ids[a];
    hb[a].state_offset = state_offsets[a];
    hb[a].control_offset = control_offsets[a];
    for (int i = 0; i < kMaxParams; ++i) hb[a].params[i] = params[static_cast<size_t>(a) * kMaxParams + i];
  }
  EvalBlock* d_b = nullptr;
  double* d_buf = nullptr;
  const size_t nd = static_cast<size_t>(2) * total_x + total_u + 2 * n_blocks + 2;
  MAS_CUDA_CHECK(cudaSetDevice(ctx->device));
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_b), n_blocks * sizeof(EvalBlock)));
  if (cudaMalloc(reinterpret_cast<void**>(&d_buf), nd * sizeof(double)) != cudaSuccess) {
    cudaFree(d_b);
    set_last_error("cudaMalloc failed");
    return MAS_B200_ERR_CUDA;
  }
  double *dX = d_buf, *dU = dX + total_x, *dD = dU + total_u, *dT = dD + total_x, *dS = dT + 2 * n_blocks;
  cudaStream_t st = ctx->stream;
  int rc = MAS_B200_OK;
  auto run = [&]() -> int {
    MAS_CUDA_CHECK(cudaMemcpyAsync(d_b, hb.data(), n_blocks * sizeof(EvalBlock), cudaMemcpyHostToDevice, st));
    MAS_CUDA_CHECK(cudaMemcpyAsync(dX, X, total_x * sizeof(double), cudaMemcpyHostToDevice, st));
    MAS_CUDA_CHECK(cudaMemcpyAsync(dU, U, total_u * sizeof(double), cudaMemcpyHostToDevice, st));
    mixed_global_eval_kernel<<<div_up(n_blocks, 64), 64, 0, st>>>(d_b, n_blocks, dX, dU, t, dD, dT, dS);
    mixed_global_sum_kernel<<<1, 1, 0, st>>>(dT, n_blocks, dS);
    MAS_CUDA_CHECK(cudaGetLastError());
    double sums[2];
    MAS_CUDA_CHECK(cudaMemcpyAsync(dyn_out, dD, total_x * sizeof(double), cudaMemcpyDeviceToHost, st));
    MAS_CUDA_CHECK(cudaMemcpyAsync(sums, dS, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    MAS_CUDA_CHECK(cudaStreamSynchronize(st));
    *stage_out = sums[0];
    *terminal_out = sums[1];
    return MAS_B200_OK;
  };
  rc = run();
  cudaFree(d_b);
  cudaFree(d_buf);
  return rc;
}

int BatchBase::nash_ls_reduce(int n_scenarios, int n_agents, int phase) {
  nash_ls_reduce_kernel<<<div_up(n_scenarios, 128), 128, 0, ctx->stream>>>(d_cost, n_scenarios, n_agents, d_base_cost, d_ls_state, phase);
  stats.kernel_launches++;
  MAS_CUDA_CHECK(cudaGetLastError());
  return MAS_B200_OK;
}

int BatchBase::upload_rows(const double* host, double* dev, int rows) {
  const size_t bytes = static_cast<size_t>(batch) * rows * sizeof(double);
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_stage, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  dim3 block(32, 8), grid(div_up(batch, 32), div_up(rows, 32));
  aos_to_soa_kernel<<<grid, block, 0, ctx->stream>>>(d_stage, dev, batch, rows, ld);
  stats.kernel_launches++;
  MAS_CUDA_CHECK(cudaGetLastError());
  return MAS_B200_OK;
}

int BatchBase::download_rows(const double* dev, double* host, int rows) {
  const size_t bytes = static_cast<size_t>(batch) * rows * sizeof(double);
  dim3 block(32, 8), grid(div_up(batch, 32), div_up(rows, 32));
  soa_to_aos_kernel<<<grid, block, 0, ctx->stream>>>(dev, d_stage, batch, rows, ld);
  stats.kernel_launches++;
  MAS_CUDA_CHECK(cudaGetLastError());
  MAS_CUDA_CHECK(cudaMemcpyAsync(host, d_stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MAS_B200_OK;
}

// Results to the host without holding up the solve stream for the PCIe transfer: transposes (HBM to HBM, fast) on
// the solve stream into a staging area of their own, the device-to-host copies on `copy_stream`.  The caller may
// start the next solve right away; wait_download() (or the next begin_download) fences the host buffers.
int BatchBase::begin_download(double* X, double* U, double* cost, int* iterations, int* status) {
  const size_t nX = static_cast<size_t>(batch) * nx * (T + 1), nU = static_cast<size_t>(batch) * nu * T;
  if (!d_out_stage) {
    MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_out_stage), (nX + nU + batch) * sizeof(double) + 2 * static_cast<size_t>(batch) * sizeof(int)));
    MAS_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_staged, cudaEventDisableTiming));
    MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_downloaded, cudaEventDisableTiming | (ctx->blocking_sync ? cudaEventBlockingSync : 0u)));
  }
  int rc = wait_download();  // the staging area and the previous host buffers are free again
  if (rc) return rc;
  double* sX = d_out_stage;
  double* sU = sX + nX;
  double* sc = sU + nU;
  int* si = reinterpret_cast<int*>(sc + batch);
  int* ss = si + batch;
  dim3 block(32, 8);
  if (X) {
    soa_to_aos_kernel<<<dim3(div_up(batch, 32), div_up(nx * (T + 1), 32)), block, 0, ctx->stream>>>(d_X, sX, batch, nx * (T + 1), ld);
    stats.kernel_launches++;
  }
  if (U) {
    soa_to_aos_kernel<<<dim3(div_up(batch, 32), div_up(nu * T, 32)), block, 0, ctx->stream>>>(d_U, sU, batch, nu * T, ld);
    stats.kernel_launches++;
  }
  MAS_CUDA_CHECK(cudaGetLastError());
  if (cost) MAS_CUDA_CHECK(cudaMemcpyAsync(sc, d_cost, batch * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (iterations) MAS_CUDA_CHECK(cudaMemcpyAsync(si, d_iters, batch * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  if (status) MAS_CUDA_CHECK(cudaMemcpyAsync(ss, d_status, batch * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
  MAS_CUDA_CHECK(cudaEventRecord(ev_staged, ctx->stream));
  MAS_CUDA_CHECK(cudaStreamWaitEvent(copy_stream, ev_staged, 0));
  if (X) MAS_CUDA_CHECK(cudaMemcpyAsync(X, sX, nX * sizeof(double), cudaMemcpyDeviceToHost, copy_stream));
  if (U) MAS_CUDA_CHECK(cudaMemcpyAsync(U, sU, nU * sizeof(double), cudaMemcpyDeviceToHost, copy_stream));
  if (cost) MAS_CUDA_CHECK(cudaMemcpyAsync(cost, sc, batch * sizeof(double), cudaMemcpyDeviceToHost, copy_stream));
  if (iterations) MAS_CUDA_CHECK(cudaMemcpyAsync(iterations, si, batch * sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
  if (status) MAS_CUDA_CHECK(cudaMemcpyAsync(status, ss, batch * sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
  MAS_CUDA_CHECK(cudaEventRecord(ev_downloaded, copy_stream));
  download_pending = true;
  return MAS_B200_OK;
}

// Scratch for the line search's trial trajectories.  Optional: when the device cannot spare it the kernels fall back to
// rolling the accepted step out a second time (same results).
int BatchBase::ensure_trial_store(long long min_slots) {
  if (trial_store_tried && trial_slots >= min_slots) return MAS_B200_OK;
  if (trial_store_tried && trial_slots == 0) return MAS_B200_OK;  // allocation failed before: stay on the fallback
  trial_store_tried = true;
  if (d_trial_X) cudaFree(d_trial_X);
  if (d_trial_U) cudaFree(d_trial_U);
  d_trial_X = d_trial_U = nullptr;
  trial_slots = 0;
  const size_t n = static_cast<size_t>(min_slots) * T;
  if (cudaMalloc(reinterpret_cast<void**>(&d_trial_X), n * nx * sizeof(double)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_trial_U), n * nu * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();  // clear the out-of-memory error; the solve proceeds without the store
    if (d_trial_X) cudaFree(d_trial_X);
    d_trial_X = d_trial_U = nullptr;
    return MAS_B200_OK;
  }
  trial_slots = static_cast<int>(min_slots);
  return MAS_B200_OK;
}

int BatchBase::wait_download() {
  if (export_pending) {  // results streamed into the sink by the last solve
    MAS_CUDA_CHECK(cudaEventSynchronize(ev_exported));
    export_pending = false;
  }
  if (!download_pending) return MAS_B200_OK;
  MAS_CUDA_CHECK(cudaEventSynchronize(ev_downloaded));
  download_pending = false;
  return MAS_B200_OK;
}

// Host buffers must be page-locked: the export kernel stores into them through their device-mapped addresses.
int BatchBase::set_result_sink(double* X, double* U, double* cost, int* iterations, int* status) {
  int rc = wait_download();
  if (rc) return rc;
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  if (export_stream) MAS_CUDA_CHECK(cudaStreamSynchronize(export_stream));
  sink_X = sink_U = sink_cost = nullptr;
  sink_iters = sink_status = nullptr;
  sink_set = false;
  if (!X && !U && !cost && !iterations && !status) return MAS_B200_OK;
  auto mapped = [&](void* host, void** dev) -> int {
    *dev = nullptr;
    if (!host) return MAS_B200_OK;
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
      cudaGetLastError();
      set_last_error("result sink: the buffers must be page-locked host memory (cudaHostAlloc / cudaHostRegister)");
      return MAS_B200_ERR_INVALID_ARGUMENT;
    }
    *dev = attr.devicePointer;
    return MAS_B200_OK;
  };
  void* d[5];
  void* h[5] = {X, U, cost, iterations, status};
  for (int i = 0; i < 5; ++i)
    if ((rc = mapped(h[i], &d[i]))) return rc;
  if (!export_stream) {
    int lo = 0, hi = 0;
    MAS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MAS_CUDA_CHECK(cudaStreamCreateWithPriority(&export_stream, cudaStreamNonBlocking, hi));  // its few CTAs should not queue behind the solves
    MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_exported, cudaEventDisableTiming | (ctx->blocking_sync ? cudaEventBlockingSync : 0u)));
    MAS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_export_src, cudaEventDisableTiming));
  }
  sink_X = static_cast<double*>(d[0]);
  sink_U = static_cast<double*>(d[1]);
  sink_cost = static_cast<double*>(d[2]);
  sink_iters = static_cast<int*>(d[3]);
  sink_status = static_cast<int*>(d[4]);
  sink_set = true;
  return MAS_B200_OK;
}

int BatchBase::export_results(int mode, int finished_at, int max_iterations, cudaEvent_t after) {
  if (!after) {
    MAS_CUDA_CHECK(cudaEventRecord(ev_export_src, ctx->stream));
    after = ev_export_src;
  }
  MAS_CUDA_CHECK(cudaStreamWaitEvent(export_stream, after, 0));
  // a few CTAs: the kernel is bound by PCIe (one 256-byte store per warp instruction), not by the SMs it occupies
  static const int env_ctas = std::getenv("MAS_B200_EXPORT_CTAS") ? std::atoi(std::getenv("MAS_B200_EXPORT_CTAS")) : 0;
  const int grid = std::max(1, std::min(env_ctas > 0 ? env_ctas : kExportCtas, div_up(div_up(batch, 32), kExportBlock / 32)));
  export_results_kernel<<<grid, kExportBlock, 0, export_stream>>>(d_X, d_U, d_cost, d_iters, d_status, batch, ld, nx * (T + 1), nu * T, mode, finished_at,
                                                                  max_iterations, sink_X, sink_U, sink_cost, sink_iters, sink_status);
  stats.kernel_launches++;
  MAS_CUDA_CHECK(cudaGetLastError());
  return MAS_B200_OK;
}

int BatchBase::finish_exports() {
  MAS_CUDA_CHECK(cudaEventRecord(ev_exported, export_stream));
  export_pending = true;
  return MAS_B200_OK;
}

int BatchBase::fence_exports() {
  // the last solve's exports read X, U, cost and the flags on their own stream: whoever overwrites them waits (on the device)
  if (ev_exported && sink_set) MAS_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ev_exported, 0));
  return MAS_B200_OK;
}

void BatchBase::prof_begin(int kind) {
  if (!profiling) return;
  TimedLaunch tl{};
  for (cudaEvent_t* e : {&tl.e0, &tl.e1}) {
    if (!event_pool.empty()) {
      *e = event_pool.back();
      event_pool.pop_back();
    } else {
      cudaEventCreate(e);
    }
  }
  tl.kind = kind;
  cudaEventRecord(tl.e0, ctx->stream);
  timed.push_back(tl);
}

void BatchBase::prof_end() {
  if (!profiling || timed.empty()) return;
  cudaEventRecord(timed.back().e1, ctx->stream);
}

// Sums the per-launch durations of the solve that just ran and the exact active-problem counts.
int BatchBase::prof_collect(int trips) {
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (auto& tl : timed) {
    float ms = 0.f;
    MAS_CUDA_CHECK(cudaEventElapsedTime(&ms, tl.e0, tl.e1));
    if (tl.kind == 0) {
      profile.prologue_ms += ms;
      profile.prologue_launches++;
    } else if (tl.kind == 1) {
      profile.backward_ms += ms;
      profile.backward_launches++;
    } else {
      profile.forward_ms += ms;
      profile.forward_launches++;
    }
    event_pool.push_back(tl.e0);
    event_pool.push_back(tl.e1);
  }
  timed.clear();
  if (trips > 0) {
    std::vector<int> hist(trips);
    MAS_CUDA_CHECK(cudaMemcpy(hist.data(), d_count_hist, trips * sizeof(int), cudaMemcpyDeviceToHost));
    for (int c : hist) profile.problem_iterations += c;
  }
  profile.solves++;
  return MAS_B200_OK;
}

int BatchBase::collect_stats() {
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  std::vector<int> it(batch), tr(batch), rg(batch);
  MAS_CUDA_CHECK(cudaMemcpy(it.data(), d_iters, batch * sizeof(int), cudaMemcpyDeviceToHost));
  MAS_CUDA_CHECK(cudaMemcpy(tr.data(), d_trials, batch * sizeof(int), cudaMemcpyDeviceToHost));
  MAS_CUDA_CHECK(cudaMemcpy(rg.data(), d_reg, batch * sizeof(int), cudaMemcpyDeviceToHost));
  long long a = 0, b = 0, c = 0;
  for (int i = 0; i < batch; ++i) {
    a += it[i];
    b += tr[i];
    c += rg[i];
  }
  stats.iterations = a;
  stats.alpha_trials = b;
  stats.reg_retries = c;
  return MAS_B200_OK;
}

}  // namespace mas_b200
