// stacked_mixed_host.cuh -- kernel wrapper and host launcher of the centralized solve over agents of different models.
#pragma once
#include <vector>

#include "engine.cuh"
#include "stacked_mixed.cuh"

namespace mas_b200 {

constexpr int kMixedThreads = 128;

// One CTA per scenario (grid-stride when there are more scenarios than resident CTAs); the workspace belongs to the CTA.
__global__ void __launch_bounds__(kMixedThreads) centralized_mixed_kernel(MixedStacked base, int n_scenarios, size_t work_stride) {
  for (int s = blockIdx.x; s < n_scenarios; s += gridDim.x) {
    MixedStacked P = base;
    P.x0 = base.x0 + static_cast<size_t>(s) * base.ns;
    P.prm = base.prm + static_cast<size_t>(s) * base.n_blocks * kMaxParams;
    P.X = base.X + static_cast<size_t>(s) * (base.T + 1) * base.ns;
    P.U = base.U + static_cast<size_t>(s) * base.T * base.ms;
    P.out_cost = base.out_cost + static_cast<size_t>(s) * (1 + base.n_blocks);
    P.out_int = base.out_int + static_cast<size_t>(s) * 4;
    P.work = base.work + static_cast<size_t>(blockIdx.x) * work_stride;
    mixed_stacked_solve(P, threadIdx.x, blockDim.x);
    __syncthreads();
  }
}

// CentralizedStrategy::operator() on n_scenarios scenarios of the same block structure.  blocks: block order, offsets and
// shapes filled in; x0 [S][ns]; params [S][n_blocks][kMaxParams]; lo / hi [ms] (has_bounds only).  Results in the stacked
// layout: X [S][T+1][ns], U [S][T][ms], costs [S][1 + n_blocks] (stacked best_cost, then each block's own), ints [S][4]
// (iterations, status, regularisation retries, line-search candidates).
inline int run_centralized_mixed(Context* ctx, const std::vector<MixedBlock>& blocks, int ns, int ms, int T, double dt, int has_bounds, const double* lo,
                                 const double* hi, const mas_b200_ilqr_params& prm, int S, const double* x0, const double* params, double* X, double* U,
                                 double* costs, int* ints, long long* launches) {
  const int nb = static_cast<int>(blocks.size());
  const MixedWork W(ns, ms, T, nb);
  cudaStream_t st = ctx->stream;
  MAS_CUDA_CHECK(cudaSetDevice(ctx->device));
  std::vector<int> box(ns), bou(ms);
  for (int a = 0; a < nb; ++a) {
    for (int i = 0; i < blocks[a].nx; ++i) box[blocks[a].state_offset + i] = a;
    for (int i = 0; i < blocks[a].nu; ++i) bou[blocks[a].control_offset + i] = a;
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, centralized_mixed_kernel, kMixedThreads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int G = std::min(S, ctx->sm_count * per_sm);
  struct Buffers {
    std::vector<void*> p;
    ~Buffers() {
      for (void* q : p) cudaFree(q);
    }
    cudaError_t get(void** out, size_t bytes) {
      const cudaError_t e = cudaMalloc(out, bytes ? bytes : 8);
      if (e == cudaSuccess) p.push_back(*out);
      return e;
    }
  } buf;
  const size_t Ss = static_cast<size_t>(S);
  MixedBlock* d_blocks = nullptr;
  int *d_box = nullptr, *d_bou = nullptr, *d_oi = nullptr;
  double *d_lo = nullptr, *d_hi = nullptr, *d_x0 = nullptr, *d_prm = nullptr, *d_X = nullptr, *d_U = nullptr, *d_work = nullptr, *d_oc = nullptr;
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_blocks), nb * sizeof(MixedBlock)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_box), ns * sizeof(int)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_bou), ms * sizeof(int)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_lo), ms * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_hi), ms * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_x0), Ss * ns * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_prm), Ss * nb * kMaxParams * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_X), Ss * (T + 1) * ns * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_U), Ss * T * ms * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_work), static_cast<size_t>(G) * W.total * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_oc), Ss * (1 + nb) * sizeof(double)));
  MAS_CUDA_CHECK(buf.get(reinterpret_cast<void**>(&d_oi), Ss * 4 * sizeof(int)));
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_blocks, blocks.data(), nb * sizeof(MixedBlock), cudaMemcpyHostToDevice, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_box, box.data(), ns * sizeof(int), cudaMemcpyHostToDevice, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_bou, bou.data(), ms * sizeof(int), cudaMemcpyHostToDevice, st));
  if (has_bounds) {
    MAS_CUDA_CHECK(cudaMemcpyAsync(d_lo, lo, ms * sizeof(double), cudaMemcpyHostToDevice, st));
    MAS_CUDA_CHECK(cudaMemcpyAsync(d_hi, hi, ms * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_x0, x0, Ss * ns * sizeof(double), cudaMemcpyHostToDevice, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(d_prm, params, Ss * nb * kMaxParams * sizeof(double), cudaMemcpyHostToDevice, st));
  MixedStacked P{};
  P.n_blocks = nb;
  P.ns = ns;
  P.ms = ms;
  P.T = T;
  P.dt = dt;
  P.has_bounds = has_bounds;
  P.tolerance = prm.tolerance;
  P.max_iterations = prm.max_iterations;
  P.max_ms = prm.max_ms;
  P.blocks = d_blocks;
  P.block_of_x = d_box;
  P.block_of_u = d_bou;
  P.lo = d_lo;
  P.hi = d_hi;
  P.x0 = d_x0;
  P.prm = d_prm;
  P.X = d_X;
  P.U = d_U;
  P.work = d_work;
  P.out_cost = d_oc;
  P.out_int = d_oi;
  centralized_mixed_kernel<<<G, kMixedThreads, 0, st>>>(P, S, W.total);
  if (launches) (*launches)++;
  MAS_CUDA_CHECK(cudaGetLastError());
  MAS_CUDA_CHECK(cudaMemcpyAsync(X, d_X, Ss * (T + 1) * ns * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(U, d_U, Ss * T * ms * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(costs, d_oc, Ss * (1 + nb) * sizeof(double), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaMemcpyAsync(ints, d_oi, Ss * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  MAS_CUDA_CHECK(cudaStreamSynchronize(st));
  return MAS_B200_OK;
}

}  // namespace mas_b200
