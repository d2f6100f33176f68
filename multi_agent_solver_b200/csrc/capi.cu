// capi.cu -- extern "C" entry points declared in include/mas_b200.h.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <limits>
#include <new>
#include <random>

#include "centralized_host.cuh"
#include "engine.cuh"

namespace mas_b200 {
const std::string& last_error();

struct ModelInfo {
  int nx, nu, np;
  unsigned available, example_mask;
  double default_params[kMaxParams];
  BatchBase* (*make)();
};

static const ModelInfo kModels[MAS_B200_NUM_MODELS] = {
    {StLane::NX, StLane::NU, StLane::NP, StLane::AVAILABLE, StLane::EXAMPLE_MASK, {1.0, 10.0, 1.0, 0.1, 0.1}, make_batch_st_lane},
    {StCirc::NX, StCirc::NU, StCirc::NP, StCirc::AVAILABLE, StCirc::EXAMPLE_MASK, {20.0, 5.0, 1.0, 1.0, 0.001, 0.001}, make_batch_st_circ},
    {Lqr4::NX, Lqr4::NU, Lqr4::NP, Lqr4::AVAILABLE, Lqr4::EXAMPLE_MASK, {0}, make_batch_lqr4},
    {Pendulum::NX, Pendulum::NU, Pendulum::NP, Pendulum::AVAILABLE, Pendulum::EXAMPLE_MASK, {60.0}, make_batch_pendulum},
    {Rocket::NX, Rocket::NU, Rocket::NP, Rocket::AVAILABLE, Rocket::EXAMPLE_MASK, {9.81, 50.0, 5e-3, 15.0, 2.0, 0.0}, make_batch_rocket},
    {StLaneCon::NX, StLaneCon::NU, StLaneCon::NP, StLaneCon::AVAILABLE, StLaneCon::EXAMPLE_MASK, {1.0, 10.0, 1.0, 0.1, 0.1, 0.8, 0.5},
     make_batch_st_lane_con},
};

// The parameter block a description stands for: its own params, or the model's example defaults when it carries none
// (the pendulum's first parameter is its horizon: pendulum_swing_up.cpp:62-98 weights the stage cost by t / T).
static void effective_params(const mas_b200_ocp_desc& d, double* out) {
  const ModelInfo& m = kModels[d.model_id];
  for (int i = 0; i < m.np; ++i) out[i] = d.num_params ? d.params[i] : m.default_params[i];
  if (d.model_id == MAS_B200_MODEL_PENDULUM && d.num_params == 0) out[0] = static_cast<double>(d.horizon_steps);
}

static int fail(int code, const std::string& msg) {
  set_last_error(msg);
  return code;
}

// ---- NCCL, resolved at run time so the library loads without it ------------------------------------
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return MAS_B200_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(MAS_B200_ERR_NCCL, std::string("dlopen libnccl.so.2 failed: ") + dlerror());
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
  g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(h, "ncclGroupStart"));
  g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd || !g_nccl.CommDestroy)
    return fail(MAS_B200_ERR_NCCL, "libnccl is missing required symbols");
  g_nccl.handle = h;
  return MAS_B200_OK;
}

#define MAS_NCCL_CHECK(expr)                                                                                        \
  do {                                                                                                              \
    ncclResult_t r__ = (expr);                                                                                      \
    if (r__ != ncclSuccess)                                                                                         \
      return fail(MAS_B200_ERR_NCCL, std::string(#expr) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error")); \
  } while (0)

static int validate_desc(const mas_b200_ocp_desc* d) {
  if (!d) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "desc is NULL");
  if (d->model_id < 0 || d->model_id >= MAS_B200_NUM_MODELS) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown model_id");
  const ModelInfo& m = kModels[d->model_id];
  if (d->state_dim != m.nx || d->control_dim != m.nu)
    return fail(MAS_B200_ERR_INVALID_ARGUMENT, "state_dim/control_dim do not match the registered model");
  if (d->horizon_steps <= 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "horizon_steps must be positive");
  if (!(d->dt != 0.0)) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "dt is 0.0");
  if (d->deriv_mask & ~m.available) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "deriv_mask asks for an analytic derivative the model does not provide");
  if (d->num_params != 0 && d->num_params != m.np) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "num_params must be 0 or the model's parameter count");
  return MAS_B200_OK;
}

static int validate_params(const mas_b200_ilqr_params* p) {
  if (!p) return fail(MAS_B200_ERR_OUT_OF_RANGE, "solver params missing (max_iterations, tolerance, max_ms are required)");
  if (p->max_iterations < 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "max_iterations < 0");
  if (std::isnan(p->tolerance) || std::isnan(p->max_ms)) return fail(MAS_B200_ERR_OUT_OF_RANGE, "tolerance / max_ms not set");
  return MAS_B200_OK;
}

}  // namespace mas_b200

using namespace mas_b200;

struct mas_b200_batch {
  BatchBase* b;
};
struct mas_b200_context {
  Context c;
  // The batch behind the one-shot entry points (mas_b200_ilqr_solve_batch, mas_b200_strategy_run), kept for the next
  // call of the same shape: creating and destroying it allocates up to 2 GB and cost 8-19 ms on a one-problem solve
  // that itself takes 1 ms.  A call with another description or size replaces it; destroyed with the context.
  mas_b200_batch* scratch_batch = nullptr;
  mas_b200_ocp_desc scratch_desc{};
  int scratch_size = 0;
  // Per-round exchange of the Nash strategies (multi-GPU): after every outer round each rank all-gathers every
  // agent's (X, U, cost) so that all ranks hold the joint trajectory set.  The gathered set of the LAST round stays
  // here (device memory, rank-major, each rank's block in the engine's [rows][ld] layout) for
  // mas_b200_strategy_get_joint; the buffers are reused by the next run of the same shape and freed with the context.
  struct Joint {
    double *X = nullptr, *U = nullptr, *cost = nullptr;
    long long* check = nullptr;  // [world][kCheck] shape words of every rank
    size_t nX = 0, nU = 0, L = 0;
    int world = 0, batch = 0, n_scenarios = 0, n_agents = 0, valid = 0;
    double collective_ms = 0.0;
    int rounds = 0;
    void release() {
      if (X) cudaFree(X);
      if (U) cudaFree(U);
      if (cost) cudaFree(cost);
      if (check) cudaFree(check);
      X = U = cost = nullptr;
      check = nullptr;
      nX = nU = L = 0;
      world = valid = 0;
    }
  } joint;
  int agents_sharded = 0;  // 1: the ranks hold different AGENTS of the same scenarios (joint totals over all ranks)
};

namespace {
// cudaEvent pairs around the per-round collectives, destroyed on every exit path
struct EventPairs {
  std::vector<cudaEvent_t> ev;
  ~EventPairs() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
  int record(cudaStream_t st) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return 1;
    ev.push_back(e);
    return cudaEventRecord(e, st) != cudaSuccess;
  }
  double total_ms() const {
    double t = 0.0;
    for (size_t i = 0; i + 1 < ev.size(); i += 2) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) t += ms;
    }
    return t;
  }
};
}  // namespace

extern "C" {

const char* mas_b200_last_error(void) { return last_error().c_str(); }
int mas_b200_version(void) { return 100; }

void mas_b200_ilqr_default_params(mas_b200_ilqr_params* p) {
  if (!p) return;
  p->max_iterations = 50;
  p->tolerance = 1e-6;
  p->max_ms = std::numeric_limits<double>::infinity();
  p->debug = 0;
  p->penalty = 10.0;
  p->penalty_increase = 5.0;
  p->constraint_tolerance = 1e-4;
  p->inequality_activation_tolerance = 1e-6;
}

int mas_b200_model_info(int model_id, int* state_dim, int* control_dim, int* num_params, unsigned* available_mask, unsigned* example_mask,
                        double* default_params) {
  if (model_id < 0 || model_id >= MAS_B200_NUM_MODELS) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown model_id");
  const ModelInfo& m = kModels[model_id];
  if (state_dim) *state_dim = m.nx;
  if (control_dim) *control_dim = m.nu;
  if (num_params) *num_params = m.np;
  if (available_mask) *available_mask = m.available;
  if (example_mask) *example_mask = m.example_mask;
  if (default_params)
    for (int i = 0; i < m.np; ++i) default_params[i] = m.default_params[i];
  return MAS_B200_OK;
}

int mas_b200_example_desc(int model_id, mas_b200_ocp_desc* out) {
  if (model_id < 0 || model_id >= MAS_B200_NUM_MODELS || !out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown model_id or NULL desc");
  const ModelInfo& m = kModels[model_id];
  std::memset(out, 0, sizeof(*out));
  out->model_id = model_id;
  out->state_dim = m.nx;
  out->control_dim = m.nu;
  out->deriv_mask = m.example_mask;
  out->num_params = m.np;
  for (int i = 0; i < m.np; ++i) out->params[i] = m.default_params[i];
  switch (model_id) {
    case MAS_B200_MODEL_SINGLE_TRACK_LANE_CONSTRAINED:
    case MAS_B200_MODEL_SINGLE_TRACK_LANE:  // single_track_ocp.cpp:21-24,105-109
      out->horizon_steps = 80;
      out->dt = 0.1;
      out->has_input_bounds = 1;
      out->input_lower[0] = -0.7;
      out->input_lower[1] = -1.0;
      out->input_upper[0] = 0.7;
      out->input_upper[1] = 1.0;
      break;
    case MAS_B200_MODEL_SINGLE_TRACK_CIRC:  // multi_agent_single_track.cpp:36-39,66-67,108
      out->horizon_steps = 10;
      out->dt = 0.5;
      out->has_input_bounds = 1;
      out->input_lower[0] = out->input_lower[1] = -0.5;
      out->input_upper[0] = out->input_upper[1] = 0.5;
      break;
    case MAS_B200_MODEL_LQR4:  // multi_agent_lqr.cpp:108-109
      out->horizon_steps = 10;
      out->dt = 0.1;
      break;
    case MAS_B200_MODEL_PENDULUM:  // pendulum_swing_up.cpp:36-37,101-103
      out->horizon_steps = 60;
      out->dt = 0.05;
      out->has_input_bounds = 1;
      out->input_lower[0] = -5.0;
      out->input_upper[0] = 5.0;
      break;
    case MAS_B200_MODEL_ROCKET:  // rocket_max_altitude.cpp:42-43,116-120
      out->horizon_steps = 50;
      out->dt = 0.1;
      out->has_input_bounds = 1;
      out->input_lower[0] = 0.0;
      out->input_upper[0] = 20.0;
      break;
  }
  return MAS_B200_OK;
}

int mas_b200_example_controls(int model_id, int horizon_steps, double* U) {
  if (model_id < 0 || model_id >= MAS_B200_NUM_MODELS || !U || horizon_steps <= 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  const int nu = kModels[model_id].nu;
  for (int i = 0; i < horizon_steps * nu; ++i) U[i] = 0.0;
  if (model_id == MAS_B200_MODEL_PENDULUM) {  // pendulum_swing_up.cpp:109-112 (host-side setup, libm sin)
    const double torque_max = 5.0, dt = 0.05;
    for (int k = 0; k < horizon_steps; ++k) {
      const double t = k * dt;
      U[k] = 0.2 * torque_max * std::sin(2.0 * M_PI * t);
    }
  } else if (model_id == MAS_B200_MODEL_ROCKET) {  // rocket_max_altitude.cpp:131
    for (int k = 0; k < horizon_steps; ++k) U[k] = 20.0 / 2.0;
  }
  return MAS_B200_OK;
}

int mas_b200_context_create(int device_id, void* stream, mas_b200_context_t* out) {
  if (!out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  int count = 0;
  MAS_CUDA_CHECK(cudaGetDeviceCount(&count));
  if (count <= 0) return fail(MAS_B200_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device_id < 0) MAS_CUDA_CHECK(cudaGetDevice(&device_id));
  if (device_id >= count) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "device_id out of range");
  MAS_CUDA_CHECK(cudaSetDevice(device_id));
  auto* ctx = new (std::nothrow) mas_b200_context();
  if (!ctx) return fail(MAS_B200_ERR_CUDA, "out of host memory");
  ctx->c.device = device_id;
  if (stream) {
    ctx->c.stream = static_cast<cudaStream_t>(stream);
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      return fail(MAS_B200_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    ctx->c.own_stream = true;
  }
  cudaDeviceGetAttribute(&ctx->c.sm_count, cudaDevAttrMultiProcessorCount, device_id);
  *out = ctx;
  return MAS_B200_OK;
}

int mas_b200_context_destroy(mas_b200_context_t ctx) {
  if (!ctx) return MAS_B200_OK;
  cudaSetDevice(ctx->c.device);
  if (ctx->scratch_batch) mas_b200_batch_destroy(ctx->scratch_batch);
  ctx->joint.release();
  ctx->c.centralized_workspace.reset();
  if (ctx->c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(static_cast<ncclComm_t>(ctx->c.nccl_comm));
  if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
  delete ctx;
  return MAS_B200_OK;
}

int mas_b200_context_synchronize(mas_b200_context_t ctx) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
  return MAS_B200_OK;
}

int mas_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  MAS_CUDA_CHECK(cudaMallocHost(out, bytes));
  return MAS_B200_OK;
}
int mas_b200_host_free(void* p) {
  if (p) MAS_CUDA_CHECK(cudaFreeHost(p));
  return MAS_B200_OK;
}

int mas_b200_batch_create(mas_b200_context_t ctx, const mas_b200_ocp_desc* desc, int batch, mas_b200_batch_t* out) {
  if (!ctx || !out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx/out is NULL");
  int rc = validate_desc(desc);
  if (rc) return rc;
  if (batch <= 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "batch must be positive");
  const ModelInfo& m = kModels[desc->model_id];
  BatchBase* b = m.make();
  b->ctx = &ctx->c;
  b->desc = *desc;
  if (desc->num_params == 0) {
    effective_params(*desc, b->desc.params);
    b->desc.num_params = m.np;
  }
  b->batch = batch;
  b->nx = m.nx;
  b->nu = m.nu;
  b->np = m.np;
  b->T = desc->horizon_steps;
  rc = b->allocate();
  if (rc) {
    delete b;
    return rc;
  }
  *out = new mas_b200_batch{b};
  return MAS_B200_OK;
}

int mas_b200_batch_destroy(mas_b200_batch_t h) {
  if (!h) return MAS_B200_OK;
  if (h->b) {
    cudaSetDevice(h->b->ctx->device);
    cudaStreamSynchronize(h->b->ctx->stream);
    delete h->b;
  }
  delete h;
  return MAS_B200_OK;
}

#define MAS_BATCH_GUARD(h)                                                      \
  if (!(h) || !(h)->b) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "batch is NULL"); \
  BatchBase* b = (h)->b;                                                        \
  MAS_CUDA_CHECK(cudaSetDevice(b->ctx->device));

int mas_b200_batch_set_initial_states(mas_b200_batch_t h, const double* x0) {
  MAS_BATCH_GUARD(h);
  if (!x0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "x0 is NULL");
  return b->upload_rows(x0, b->d_x0, b->nx);
}

int mas_b200_batch_set_params(mas_b200_batch_t h, const double* params) {
  MAS_BATCH_GUARD(h);
  if (!params || b->np == 0) {
    b->per_problem_params = false;
    return MAS_B200_OK;
  }
  b->per_problem_params = true;
  return b->upload_rows(params, b->d_params, b->np);
}

int mas_b200_batch_set_controls(mas_b200_batch_t h, const double* U) {
  MAS_BATCH_GUARD(h);
  int frc = b->fence_exports();  // the last solve's results may still be travelling out of U
  if (frc) return frc;
  if (!U) {
    MAS_CUDA_CHECK(cudaMemsetAsync(b->d_U, 0, static_cast<size_t>(b->ld) * b->nu * b->T * sizeof(double), b->ctx->stream));
    return MAS_B200_OK;
  }
  return b->upload_rows(U, b->d_U, b->nu * b->T);
}

int mas_b200_batch_initialize(mas_b200_batch_t h) {
  MAS_BATCH_GUARD(h);
  int frc = b->fence_exports();
  if (frc) return frc;
  return b->initialize();
}

int mas_b200_batch_solve(mas_b200_batch_t h, const mas_b200_ilqr_params* params) {
  MAS_BATCH_GUARD(h);
  int rc = validate_params(params);
  if (rc) return rc;
  return b->solve(*params);
}

int mas_b200_batch_get_solution(mas_b200_batch_t h, double* X, double* U, double* cost, int* iterations, int* status) {
  MAS_BATCH_GUARD(h);
  int rc = MAS_B200_OK;
  if (X && (rc = b->download_rows(b->d_X, X, b->nx * (b->T + 1)))) return rc;
  if (U && (rc = b->download_rows(b->d_U, U, b->nu * b->T))) return rc;
  if (cost) MAS_CUDA_CHECK(cudaMemcpyAsync(cost, b->d_cost, b->batch * sizeof(double), cudaMemcpyDeviceToHost, b->ctx->stream));
  if (iterations) MAS_CUDA_CHECK(cudaMemcpyAsync(iterations, b->d_iters, b->batch * sizeof(int), cudaMemcpyDeviceToHost, b->ctx->stream));
  if (status) MAS_CUDA_CHECK(cudaMemcpyAsync(status, b->d_status, b->batch * sizeof(int), cudaMemcpyDeviceToHost, b->ctx->stream));
  MAS_CUDA_CHECK(cudaStreamSynchronize(b->ctx->stream));
  return MAS_B200_OK;
}

int mas_b200_batch_begin_get_solution(mas_b200_batch_t h, double* X, double* U, double* cost, int* iterations, int* status) {
  MAS_BATCH_GUARD(h);
  return b->begin_download(X, U, cost, iterations, status);
}

int mas_b200_batch_wait_solution(mas_b200_batch_t h) {
  MAS_BATCH_GUARD(h);
  return b->wait_download();
}

int mas_b200_batch_set_result_sink(mas_b200_batch_t h, double* X, double* U, double* cost, int* iterations, int* status) {
  MAS_BATCH_GUARD(h);
  return b->set_result_sink(X, U, cost, iterations, status);
}

int mas_b200_batch_get_device_view(mas_b200_batch_t h, mas_b200_device_view* out) {
  MAS_BATCH_GUARD(h);
  if (!out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  out->batch = b->batch;
  out->ld = b->ld;
  out->state_dim = b->nx;
  out->control_dim = b->nu;
  out->horizon_steps = b->T;
  out->x0 = b->d_x0;
  out->X = b->d_X;
  out->U = b->d_U;
  out->cost = b->d_cost;
  out->iterations = b->d_iters;
  out->status = b->d_status;
  out->params = b->d_params;
  return MAS_B200_OK;
}

int mas_b200_batch_reset_solver_state(mas_b200_batch_t h) {
  MAS_BATCH_GUARD(h);
  b->al_fresh = true;
  return MAS_B200_OK;
}

int mas_b200_batch_get_stats(mas_b200_batch_t h, mas_b200_batch_stats* out) {
  MAS_BATCH_GUARD(h);
  if (!out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  int rc = b->collect_stats();
  if (rc) return rc;
  *out = b->stats;
  return MAS_B200_OK;
}

int mas_b200_batch_set_profiling(mas_b200_batch_t h, int enable) {
  MAS_BATCH_GUARD(h);
  b->profiling = enable != 0;
  b->profile = mas_b200_profile{};
  return MAS_B200_OK;
}

int mas_b200_batch_get_profile(mas_b200_batch_t h, mas_b200_profile* out) {
  MAS_BATCH_GUARD(h);
  if (!out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = b->profile;
  return MAS_B200_OK;
}

int mas_b200_batch_set_tuning(mas_b200_batch_t h, int forward_lanes, int forward_chains) {
  MAS_BATCH_GUARD(h);
  if (forward_lanes != 0 && forward_lanes != 1 && forward_lanes != 2 && forward_lanes != 4 && forward_lanes != 8 && forward_lanes != 16)
    return fail(MAS_B200_ERR_INVALID_ARGUMENT, "forward_lanes must be 0,1,2,4,8 or 16");
  if (forward_chains < 0 || forward_chains > 2) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "forward_chains must be 0,1 or 2");
  b->tune_L = forward_lanes;
  b->tune_C = forward_chains;
  return MAS_B200_OK;
}

int mas_b200_batch_set_trial_store(mas_b200_batch_t h, int enable) {
  MAS_BATCH_GUARD(h);
  b->trial_store = enable != 0;
  b->coop_store = enable == 2;
  b->backward_lanes_enabled = enable != 0;  // the same switch turns the lane-parallel backward pass off (A/B runs, tests)
  return MAS_B200_OK;
}

int mas_b200_batch_set_backward_mode(mas_b200_batch_t h, int mode, int max_problems) {
  MAS_BATCH_GUARD(h);
  if (mode < 0 || mode > 3) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "backward mode must be 0 (auto), 1 (one thread), 2 (FD lanes) or 3 (time-parallel)");
  b->backward_mode = mode;
  if (max_problems > 0) b->tp_max_problems = max_problems;
  return MAS_B200_OK;
}

int mas_b200_batch_get_debug_trace(mas_b200_batch_t h, int problem, int max_records, double* records, int* n_records) {
  MAS_BATCH_GUARD(h);
  if (!records || !n_records || max_records < 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  if (problem < 0 || problem >= b->batch) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "problem index out of range");
  if (!b->dbg_valid || !b->d_dbg) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "no trace: the last solve of this batch did not have params.debug set");
  const int want = std::min(max_records, b->dbg_records);
  // one strided copy: `want * kDebugFields` doubles of column `problem`
  MAS_CUDA_CHECK(cudaMemcpy2DAsync(records, sizeof(double), b->d_dbg + problem, static_cast<size_t>(b->ld) * sizeof(double), sizeof(double),
                                   static_cast<size_t>(want) * kDebugFields, cudaMemcpyDeviceToHost, b->ctx->stream));
  MAS_CUDA_CHECK(cudaStreamSynchronize(b->ctx->stream));
  int n = 0;
  while (n < want && !std::isnan(records[static_cast<size_t>(n) * kDebugFields + 5 * (n > 0)])) ++n;  // unused records are all-NaN
  *n_records = n;
  return MAS_B200_OK;
}

int mas_b200_batch_set_concurrency_hint(mas_b200_batch_t h, int solves_in_flight) {
  MAS_BATCH_GUARD(h);
  if (solves_in_flight < 1) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "solves_in_flight must be >= 1");
  b->concurrency_hint = solves_in_flight;
  return MAS_B200_OK;
}

int mas_b200_batch_set_line_search_mode(mas_b200_batch_t h, int mode) {
  MAS_BATCH_GUARD(h);
  if (mode < 0 || mode > 3) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "line search mode must be 0 (auto), 1 (lanes), 2 (rounds) or 3 (warp-cooperative)");
  b->ls_mode = mode;
  return MAS_B200_OK;
}

// One-shot calls borrow the context's scratch batch; every borrow starts from a fresh solver (no multipliers, default
// tuning), as a newly constructed reference solver would.
static int borrow_scratch_batch(mas_b200_context_t ctx, const mas_b200_ocp_desc* desc, int batch, mas_b200_batch_t* out) {
  if (!ctx || !desc) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx/desc is NULL");
  if (ctx->scratch_batch && ctx->scratch_size == batch && std::memcmp(&ctx->scratch_desc, desc, sizeof(*desc)) == 0) {
    BatchBase* b = ctx->scratch_batch->b;
    b->al_fresh = true;
    b->tune_L = b->tune_C = 0;
    b->ls_mode = 0;
    b->backward_mode = 0;
    *out = ctx->scratch_batch;
    return MAS_B200_OK;
  }
  if (ctx->scratch_batch) {
    mas_b200_batch_destroy(ctx->scratch_batch);
    ctx->scratch_batch = nullptr;
  }
  mas_b200_batch_t h = nullptr;
  const int rc = mas_b200_batch_create(ctx, desc, batch, &h);
  if (rc) return rc;
  ctx->scratch_batch = h;
  std::memcpy(&ctx->scratch_desc, desc, sizeof(*desc));
  ctx->scratch_size = batch;
  *out = h;
  return MAS_B200_OK;
}

int mas_b200_ilqr_solve_batch(mas_b200_context_t ctx, const mas_b200_ocp_desc* desc, const mas_b200_ilqr_params* params, int batch,
                              const double* x0, const double* model_params, double* U, double* X, double* cost, int* iterations, int* status) {
  int rc = validate_params(params);
  if (rc) return rc;
  mas_b200_batch_t h = nullptr;
  rc = borrow_scratch_batch(ctx, desc, batch, &h);
  if (rc) return rc;
  rc = mas_b200_batch_set_initial_states(h, x0);
  if (!rc) rc = mas_b200_batch_set_params(h, model_params);
  if (!rc) rc = mas_b200_batch_set_controls(h, U);
  if (!rc) rc = mas_b200_batch_solve(h, params);
  if (!rc) rc = mas_b200_batch_get_solution(h, X, U, cost, iterations, status);
  return rc;
}

int mas_b200_ilqr_last_debug_trace(mas_b200_context_t ctx, int problem, int max_records, double* records, int* n_records) {
  if (!ctx || !ctx->scratch_batch) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "no one-shot solve has run on this context");
  return mas_b200_batch_get_debug_trace(ctx->scratch_batch, problem, max_records, records, n_records);
}

// ---- strategies ------------------------------------------------------------------------------------------
// CentralizedStrategy through the general stacked solve (stacked_mixed.cuh: run-time block shapes, workspace in HBM): agents of
// different models, and stacks of one model that are too large for the compiled-in kernel of centralized.cuh.  Per-agent
// pointer arrays as in mas_b200_strategy_run_mixed.
static int run_centralized_general(mas_b200_context_t ctx, const mas_b200_ocp_desc* agent_descs, const mas_b200_ilqr_params* params, int max_outer,
                                   int n_scenarios, int n_agents, const double* const* x0, const double* const* model_params, double* const* X,
                                   double* const* U, double* const* costs, double* total_cost, int* trace_iterations) {
  int rc = MAS_B200_OK;
  // build_global_ocp of a mixed problem (multi_agent_problem.hpp:52-127): blocks in agent order (ids 0..n-1), horizon and dt of the
  // first block, bounds only when every agent has both; the stacked solve starts from zero controls (U_init is not read)
  const int S = n_scenarios, T = agent_descs[0].horizon_steps;
  std::vector<int> mid(n_agents), soff(n_agents), uoff(n_agents);
  int ns = 0, ms = 0;
  bool all_bounds = true;
  for (int a = 0; a < n_agents; ++a) {
    const mas_b200_ocp_desc& d = agent_descs[a];
    mid[a] = d.model_id;
    soff[a] = ns;
    uoff[a] = ms;
    ns += d.state_dim;
    ms += d.control_dim;
    all_bounds = all_bounds && d.has_input_bounds;
  }
  std::vector<double> lo(ms, 0.0), hi(ms, 0.0), fx0(static_cast<size_t>(S) * ns), fp(static_cast<size_t>(S) * n_agents * kMaxParams, 0.0);
  for (int a = 0; a < n_agents; ++a) {
    const mas_b200_ocp_desc& d = agent_descs[a];
    const int np = kModels[d.model_id].np;
    for (int i = 0; i < d.control_dim; ++i) {
      lo[uoff[a] + i] = d.input_lower[i];
      hi[uoff[a] + i] = d.input_upper[i];
    }
    for (int sc = 0; sc < S; ++sc) {
      std::memcpy(&fx0[static_cast<size_t>(sc) * ns + soff[a]], x0[a] + static_cast<size_t>(sc) * d.state_dim, sizeof(double) * d.state_dim);
      double* dst = &fp[(static_cast<size_t>(sc) * n_agents + a) * kMaxParams];
      if (model_params && model_params[a])
        for (int i = 0; i < np; ++i) dst[i] = model_params[a][static_cast<size_t>(sc) * np + i];
      else
        effective_params(d, dst);
    }
  }
  std::vector<double> gX(static_cast<size_t>(S) * (T + 1) * ns), gU(static_cast<size_t>(S) * T * ms), gc(static_cast<size_t>(S) * (1 + n_agents));
  std::vector<int> gi(static_cast<size_t>(S) * 4);
  rc = centralized_mixed_entry(&ctx->c, n_agents, mid.data(), T, agent_descs[0].dt, all_bounds ? 1 : 0, lo.data(), hi.data(), *params, S, fx0.data(),
                               fp.data(), gX.data(), gU.data(), gc.data(), gi.data(), nullptr);
  if (rc) return rc;
  // every agent gets its rows of the stacked result (centralized.hpp:27-36): shapes [scenario][T+1][n_a] / [scenario][T][m_a], T of the first block
  for (int sc = 0; sc < S; ++sc) {
    for (int a = 0; a < n_agents; ++a) {
      const int n = agent_descs[a].state_dim, m = agent_descs[a].control_dim;
      if (X && X[a])
        for (int t = 0; t <= T; ++t)
          std::memcpy(X[a] + (static_cast<size_t>(sc) * (T + 1) + t) * n, &gX[(static_cast<size_t>(sc) * (T + 1) + t) * ns + soff[a]], sizeof(double) * n);
      if (U && U[a])
        for (int t = 0; t < T; ++t)
          std::memcpy(U[a] + (static_cast<size_t>(sc) * T + t) * m, &gU[(static_cast<size_t>(sc) * T + t) * ms + uoff[a]], sizeof(double) * m);
      if (costs && costs[a]) costs[a][sc] = gc[static_cast<size_t>(sc) * (1 + n_agents) + 1 + a];
    }
    if (total_cost) total_cost[sc] = gc[static_cast<size_t>(sc) * (1 + n_agents)];
    if (trace_iterations && max_outer >= 1) trace_iterations[static_cast<size_t>(sc) * max_outer * n_agents] = gi[static_cast<size_t>(sc) * 4];
  }
  return MAS_B200_OK;
}

int mas_b200_strategy_run(mas_b200_context_t ctx, int strategy, const mas_b200_ocp_desc* agent_desc, const mas_b200_ilqr_params* params,
                          int max_outer, int n_scenarios, int n_agents, const double* x0, const double* model_params, const double* U_init,
                          double* X, double* U, double* costs, double* total_cost, int* trace_iterations, int* trace_accepted, double* trace_cost) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  int rc = validate_params(params);
  if (rc) return rc;
  if (n_scenarios <= 0 || n_agents <= 0 || max_outer < 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad scenario/agent/outer counts");
  if (strategy == MAS_B200_STRATEGY_CENTRALIZED) {
    // stack -> one solve -> scatter (strategies/centralized.hpp:18-38); the stacked problem is always all-FD
    rc = validate_desc(agent_desc);
    if (rc) return rc;
    if (n_agents * agent_desc->state_dim > 256) {
      // the kernel of centralized.cuh keeps its index arithmetic and scratch sized for stacked states up to 256; larger stacks
      // take the general solve with its workspace in HBM (bit-identical results, slower): re-pack [scenario][agent][...]
      // into the per-agent arrays it works on
      const int S = n_scenarios, A = n_agents, n = agent_desc->state_dim, m = agent_desc->control_dim, T = agent_desc->horizon_steps;
      const int np = kModels[agent_desc->model_id].np;
      const size_t Ss = static_cast<size_t>(S);
      std::vector<mas_b200_ocp_desc> descs(A, *agent_desc);
      std::vector<double> ax0(Ss * A * n), ap(model_params ? Ss * A * np : 0), aX(Ss * A * (T + 1) * n), aU(Ss * A * T * m), ac(Ss * A);
      std::vector<const double*> px0(A), pp(A, nullptr);
      std::vector<double*> pX(A), pU(A), pc(A);
      for (int a = 0; a < A; ++a) {
        px0[a] = &ax0[Ss * a * n];
        if (model_params) pp[a] = &ap[Ss * a * np];
        pX[a] = &aX[Ss * a * (T + 1) * n];
        pU[a] = &aU[Ss * a * T * m];
        pc[a] = &ac[Ss * a];
        for (int sc = 0; sc < S; ++sc) {
          std::memcpy(&ax0[(Ss * a + sc) * n], x0 + (static_cast<size_t>(sc) * A + a) * n, sizeof(double) * n);
          if (model_params) std::memcpy(&ap[(Ss * a + sc) * np], model_params + (static_cast<size_t>(sc) * A + a) * np, sizeof(double) * np);
        }
      }
      rc = run_centralized_general(ctx, descs.data(), params, max_outer, S, A, px0.data(), model_params ? pp.data() : nullptr, pX.data(), pU.data(),
                                   pc.data(), total_cost, trace_iterations);  // [scenario][outer][agent] in both layouts
      if (rc) return rc;
      for (int sc = 0; sc < S; ++sc)
        for (int a = 0; a < A; ++a) {
          const size_t sa = static_cast<size_t>(sc) * A + a;
          if (X) std::memcpy(X + sa * (T + 1) * n, &aX[(Ss * a + sc) * (T + 1) * n], sizeof(double) * (T + 1) * n);
          if (U) std::memcpy(U + sa * T * m, &aU[(Ss * a + sc) * T * m], sizeof(double) * T * m);
          if (costs) costs[sa] = ac[Ss * a + sc];
        }
      return MAS_B200_OK;
    }
    CentralizedFn fn = centralized_entry(agent_desc->model_id);
    if (!fn) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown model_id");
    mas_b200_ocp_desc d = *agent_desc;
    if (d.num_params == 0) {
      effective_params(*agent_desc, d.params);
      d.num_params = kModels[d.model_id].np;
    }
    std::vector<int> its(n_scenarios);
    rc = fn(&ctx->c, d, *params, n_scenarios, n_agents, x0, model_params, U_init, X, U, costs, total_cost, its.data(), nullptr, nullptr);
    if (rc) return rc;
    // the documented trace shape is [scenario][max_outer][agent]: with max_outer == 0 it has no elements, so nothing is
    // written (max_outer has no meaning for the centralized strategy; pass 1 to receive the iteration count)
    if (trace_iterations && max_outer >= 1)
      for (int s = 0; s < n_scenarios; ++s) trace_iterations[static_cast<size_t>(s) * max_outer * n_agents] = its[s];
    return MAS_B200_OK;
  }
  if (strategy != MAS_B200_STRATEGY_SEQUENTIAL && strategy != MAS_B200_STRATEGY_TRUSTREGION && strategy != MAS_B200_STRATEGY_LINESEARCH)
    return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown strategy");
  const bool keeps_old = strategy == MAS_B200_STRATEGY_TRUSTREGION || strategy == MAS_B200_STRATEGY_LINESEARCH;
  const int batch = n_scenarios * n_agents;
  mas_b200_batch_t h = nullptr;
  rc = borrow_scratch_batch(ctx, agent_desc, batch, &h);
  if (rc) return rc;
  BatchBase* b = h->b;
  cudaStream_t st = ctx->c.stream;
  std::vector<int> tmp_i(batch);
  std::vector<double> tmp_c(batch);
  auto run = [&]() -> int {
    int r = mas_b200_batch_set_initial_states(h, x0);
    if (!r) r = mas_b200_batch_set_params(h, model_params);
    if (!r) r = mas_b200_batch_set_controls(h, U_init);
    if (!r) r = b->initialize();  // OCP::initialize_problem of every agent
    if (r) return r;
    const size_t L = static_cast<size_t>(b->ld);
    if (keeps_old) {
      r = b->ensure_strategy_scratch();
      if (r) return r;
    }
    if (strategy == MAS_B200_STRATEGY_LINESEARCH) {
      r = b->nash_ls_reduce(n_scenarios, n_agents, -1);  // base_cost = total_cost(problem), nash.hpp:103
      if (r) return r;
    }
    if (strategy == MAS_B200_STRATEGY_TRUSTREGION) {
      std::vector<double> ones(b->ld, 1.0);  // radii = 1.0 (nash.hpp:194)
      MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_radius, ones.data(), L * sizeof(double), cudaMemcpyHostToDevice, st));
      MAS_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    // joint buffers for the per-round exchange of every agent's (X, U, cost) across ranks (kept in the context)
    const size_t nX = L * b->nx * (b->T + 1), nU = L * b->nu * b->T;
    mas_b200_context::Joint& J = ctx->joint;
    EventPairs coll_events;
    J.valid = 0;
    if (ctx->c.nccl_comm) {
      const int world = ctx->c.world;
      if (J.world != world || J.nX != nX || J.nU != nU || J.L != L) {
        J.release();
        MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&J.X), nX * world * sizeof(double)));
        MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&J.U), nU * world * sizeof(double)));
        MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&J.cost), L * world * sizeof(double)));
        MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&J.check), static_cast<size_t>(world + 1) * 8 * sizeof(long long)));
        J.world = world;
        J.nX = nX;
        J.nU = nU;
        J.L = L;
      }
      J.batch = batch;
      J.n_scenarios = n_scenarios;
      J.n_agents = n_agents;
      // every rank must bring the same shape to the collectives (equal counts are what ncclAllGather assumes): exchange
      // the shape words first and fail on ALL ranks alike if they differ, instead of hanging or corrupting memory
      const long long mine[8] = {static_cast<long long>(L), batch, max_outer, n_agents, static_cast<long long>(nX), static_cast<long long>(nU),
                                 strategy, n_scenarios};
      ncclComm_t comm = static_cast<ncclComm_t>(ctx->c.nccl_comm);
      MAS_CUDA_CHECK(cudaMemcpyAsync(J.check + static_cast<size_t>(world) * 8, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
      MAS_NCCL_CHECK(g_nccl.AllGather(J.check + static_cast<size_t>(world) * 8, J.check, 8, ncclInt64, comm, st));
      std::vector<long long> all(static_cast<size_t>(world) * 8);
      MAS_CUDA_CHECK(cudaMemcpyAsync(all.data(), J.check, all.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
      MAS_CUDA_CHECK(cudaStreamSynchronize(st));
      for (int rk = 0; rk < world; ++rk)
        for (int k = 0; k < 8; ++k)
          if (all[static_cast<size_t>(rk) * 8 + k] != mine[k])
            return fail(MAS_B200_ERR_INVALID_ARGUMENT,
                        "multi-GPU strategy run: ranks disagree on the local shape (scenarios x agents, horizon, max_outer or strategy); "
                        "give every rank the same number of scenarios and agents (pad the last shard)");
    }
    // per-round record: inner iteration counts, accept flags, costs ([scenario][outer][agent])
    auto record_trace = [&](int outer) -> int {
      if (!(trace_iterations || trace_accepted || trace_cost)) return MAS_B200_OK;
      const size_t off = static_cast<size_t>(outer) * n_agents;
      if (trace_iterations) {
        MAS_CUDA_CHECK(cudaMemcpyAsync(tmp_i.data(), b->d_iters, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
        MAS_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int s = 0; s < n_scenarios; ++s)
          for (int a = 0; a < n_agents; ++a) trace_iterations[static_cast<size_t>(s) * max_outer * n_agents + off + a] = tmp_i[s * n_agents + a];
      }
      if (trace_accepted) {
        if (strategy == MAS_B200_STRATEGY_TRUSTREGION) {
          MAS_CUDA_CHECK(cudaMemcpyAsync(tmp_i.data(), b->d_accepted, batch * sizeof(int), cudaMemcpyDeviceToHost, st));
          MAS_CUDA_CHECK(cudaStreamSynchronize(st));
        } else {
          for (auto& v : tmp_i) v = 1;
        }
        for (int s = 0; s < n_scenarios; ++s)
          for (int a = 0; a < n_agents; ++a) trace_accepted[static_cast<size_t>(s) * max_outer * n_agents + off + a] = tmp_i[s * n_agents + a];
      }
      if (trace_cost) {
        MAS_CUDA_CHECK(cudaMemcpyAsync(tmp_c.data(), b->d_cost, batch * sizeof(double), cudaMemcpyDeviceToHost, st));
        MAS_CUDA_CHECK(cudaStreamSynchronize(st));
        for (int s = 0; s < n_scenarios; ++s)
          for (int a = 0; a < n_agents; ++a) trace_cost[static_cast<size_t>(s) * max_outer * n_agents + off + a] = tmp_c[s * n_agents + a];
      }
      return MAS_B200_OK;
    };
    for (int outer = 0; outer < max_outer; ++outer) {
      if (keeps_old) {  // nash.hpp:108-115 / :208-210
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_U_old, b->d_U, nU * sizeof(double), cudaMemcpyDeviceToDevice, st));
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_X_old, b->d_X, nX * sizeof(double), cudaMemcpyDeviceToDevice, st));
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_cost_old, b->d_cost, L * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
      r = b->solve(*params);  // nash.hpp:59-64 / :212, one warm-started solve per agent
      if (r) return r;
      if (strategy == MAS_B200_STRATEGY_TRUSTREGION) {
        r = b->trust_region_step();  // nash.hpp:218-243
        if (r) return r;
      }
      if (strategy == MAS_B200_STRATEGY_LINESEARCH) {
        r = record_trace(outer);  // the reference's per-solve record sits inside sequential_solve, before the joint search
        if (r) return r;
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_U_cand, b->d_U, nU * sizeof(double), cudaMemcpyDeviceToDevice, st));  // nash.hpp:123-125
        r = b->nash_ls_reduce(n_scenarios, n_agents, 0);  // new_cost >= base_cost ? search : base_cost = new_cost
        if (r) return r;
        for (double alpha = 0.5; alpha > 1e-3; alpha *= 0.5) {  // nash.hpp:127-158
          r = b->nash_ls_trial(n_agents, alpha);
          if (!r) r = b->nash_ls_reduce(n_scenarios, n_agents, 1);
          if (r) return r;
        }
        r = b->nash_ls_restore(n_agents);  // nash.hpp:161-171
        if (r) return r;
      }
      if (ctx->c.nccl_comm) {  // every rank ends the round holding all agents' trajectories
        ncclComm_t comm = static_cast<ncclComm_t>(ctx->c.nccl_comm);
        if (coll_events.record(st)) return fail(MAS_B200_ERR_CUDA, "cudaEventRecord failed");
        MAS_NCCL_CHECK(g_nccl.GroupStart());
        MAS_NCCL_CHECK(g_nccl.AllGather(b->d_X, J.X, nX, ncclDouble, comm, st));
        MAS_NCCL_CHECK(g_nccl.AllGather(b->d_U, J.U, nU, ncclDouble, comm, st));
        MAS_NCCL_CHECK(g_nccl.AllGather(b->d_cost, J.cost, L, ncclDouble, comm, st));
        MAS_NCCL_CHECK(g_nccl.GroupEnd());
        if (coll_events.record(st)) return fail(MAS_B200_ERR_CUDA, "cudaEventRecord failed");
      }
      if (strategy != MAS_B200_STRATEGY_LINESEARCH) {
        r = record_trace(outer);
        if (r) return r;
      }
    }
    // collect_solution (nash.hpp:23-37): per-agent trajectories and costs, total in block order
    std::vector<double> c(batch);
    r = mas_b200_batch_get_solution(h, X, U, c.data(), nullptr, nullptr);
    if (r) return r;
    if (ctx->c.nccl_comm) {
      J.collective_ms = coll_events.total_ms();
      J.rounds = max_outer;
      J.valid = max_outer > 0 ? 1 : 0;
    }
    if (ctx->c.nccl_comm && ctx->agents_sharded && max_outer > 0) {
      // the ranks hold different agents of the same scenarios: the scenario total runs over ALL ranks' agents, in block
      // order (rank-major = id order), from the gathered costs of the last round -- the same sum on every rank
      std::vector<double> all(static_cast<size_t>(ctx->c.world) * L);
      MAS_CUDA_CHECK(cudaMemcpyAsync(all.data(), J.cost, all.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
      MAS_CUDA_CHECK(cudaStreamSynchronize(st));
      for (int s = 0; s < n_scenarios; ++s) {
        double tot = 0.0;
        for (int rk = 0; rk < ctx->c.world; ++rk)
          for (int a = 0; a < n_agents; ++a) tot += all[static_cast<size_t>(rk) * L + static_cast<size_t>(s) * n_agents + a];
        if (total_cost) total_cost[s] = tot;
      }
    } else {
      for (int s = 0; s < n_scenarios; ++s) {
        double tot = 0.0;
        for (int a = 0; a < n_agents; ++a) tot += c[static_cast<size_t>(s) * n_agents + a];
        if (total_cost) total_cost[s] = tot;
      }
    }
    if (costs) std::memcpy(costs, c.data(), sizeof(double) * batch);
    return MAS_B200_OK;
  };
  rc = run();  // the batch stays with the context (borrow_scratch_batch)
  return rc;
}

// ---- strategies over agents of different models / shapes ------------------------------------------------------------
namespace {
struct AgentGroup {  // agents that share one description: one device batch [scenario][member]
  mas_b200_ocp_desc desc;
  std::vector<int> members;  // agent indices, ascending
  mas_b200_batch* h = nullptr;
  std::vector<double> cost;  // [scenario][member] after the last download
};
struct GroupSet {
  std::vector<AgentGroup> g;
  ~GroupSet() {
    for (auto& a : g)
      if (a.h) mas_b200_batch_destroy(a.h);
  }
};
}  // namespace

int mas_b200_strategy_run_mixed(mas_b200_context_t ctx, int strategy, const mas_b200_ocp_desc* agent_descs, const mas_b200_ilqr_params* params,
                                int max_outer, int n_scenarios, int n_agents, const double* const* x0, const double* const* model_params,
                                const double* const* U_init, double* const* X, double* const* U, double* const* costs, double* total_cost,
                                int* trace_iterations) {
  if (!ctx || !agent_descs || !x0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx / agent_descs / x0 is NULL");
  int rc = validate_params(params);
  if (rc) return rc;
  if (n_scenarios <= 0 || n_agents <= 0 || max_outer < 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad scenario/agent/outer counts");
  for (int a = 0; a < n_agents; ++a) {
    rc = validate_desc(&agent_descs[a]);
    if (rc) return rc;
  }
  bool same = true;
  for (int a = 1; a < n_agents; ++a) same = same && std::memcmp(&agent_descs[a], &agent_descs[0], sizeof(mas_b200_ocp_desc)) == 0;
  if (strategy == MAS_B200_STRATEGY_CENTRALIZED && !same)
    return run_centralized_general(ctx, agent_descs, params, max_outer, n_scenarios, n_agents, x0, model_params, X, U, costs, total_cost, trace_iterations);
  if (strategy != MAS_B200_STRATEGY_CENTRALIZED && strategy != MAS_B200_STRATEGY_SEQUENTIAL && strategy != MAS_B200_STRATEGY_TRUSTREGION &&
      strategy != MAS_B200_STRATEGY_LINESEARCH)
    return fail(MAS_B200_ERR_INVALID_ARGUMENT, "unknown strategy");
  const int S = n_scenarios;
  if (same) {  // one shape: pack the per-agent arrays into the [scenario][agent] layout and take the single-batch path
    const mas_b200_ocp_desc& d = agent_descs[0];
    const int n = d.state_dim, m = d.control_dim, T = d.horizon_steps, np = kModels[d.model_id].np;
    const size_t SA = static_cast<size_t>(S) * n_agents;
    std::vector<double> fx0(SA * n), fp, fU0, fX(SA * n * (T + 1)), fU(SA * m * T), fc(SA);
    bool any_p = false, any_u = false;
    for (int a = 0; a < n_agents; ++a) {
      any_p = any_p || (model_params && model_params[a]);
      any_u = any_u || (U_init && U_init[a]);
    }
    if (any_p) fp.assign(SA * np, 0.0);
    if (any_u) fU0.assign(SA * m * T, 0.0);
    for (int sc = 0; sc < S; ++sc)
      for (int a = 0; a < n_agents; ++a) {
        const size_t idx = static_cast<size_t>(sc) * n_agents + a;
        std::memcpy(&fx0[idx * n], x0[a] + static_cast<size_t>(sc) * n, sizeof(double) * n);
        if (any_p) {
          if (model_params[a])
            for (int i = 0; i < np; ++i) fp[idx * np + i] = model_params[a][static_cast<size_t>(sc) * np + i];
          else
            effective_params(d, &fp[idx * np]);
        }
        if (any_u && U_init[a]) std::memcpy(&fU0[idx * m * T], U_init[a] + static_cast<size_t>(sc) * m * T, sizeof(double) * m * T);
      }
    rc = mas_b200_strategy_run(ctx, strategy, &d, params, max_outer, S, n_agents, fx0.data(), any_p ? fp.data() : nullptr, any_u ? fU0.data() : nullptr,
                               fX.data(), fU.data(), fc.data(), total_cost, trace_iterations, nullptr, nullptr);
    if (rc) return rc;
    for (int sc = 0; sc < S; ++sc)
      for (int a = 0; a < n_agents; ++a) {
        const size_t idx = static_cast<size_t>(sc) * n_agents + a;
        if (X && X[a]) std::memcpy(X[a] + static_cast<size_t>(sc) * n * (T + 1), &fX[idx * n * (T + 1)], sizeof(double) * n * (T + 1));
        if (U && U[a]) std::memcpy(U[a] + static_cast<size_t>(sc) * m * T, &fU[idx * m * T], sizeof(double) * m * T);
        if (costs && costs[a]) costs[a][sc] = fc[idx];
      }
    return MAS_B200_OK;
  }
  // ---- Nash strategies, mixed agents: agents never read each other's trajectories during a solve (nash.hpp:59-64,199-212),
  // so agents of one description form one device batch and the groups advance round by round in lockstep; the only joint
  // quantity, the line-search strategy's total cost (nash.hpp:39-51,103,121,143), is summed on the host in block order.
  GroupSet gs;
  for (int a = 0; a < n_agents; ++a) {
    AgentGroup* found = nullptr;
    for (auto& g : gs.g)
      if (std::memcmp(&g.desc, &agent_descs[a], sizeof(mas_b200_ocp_desc)) == 0) found = &g;
    if (!found) {
      gs.g.push_back(AgentGroup{});
      found = &gs.g.back();
      found->desc = agent_descs[a];
    }
    found->members.push_back(a);
  }
  cudaStream_t st = ctx->c.stream;
  const bool keeps_old = strategy != MAS_B200_STRATEGY_SEQUENTIAL;
  for (auto& g : gs.g) {
    const int G = static_cast<int>(g.members.size()), B = S * G;
    const int n = g.desc.state_dim, m = g.desc.control_dim, T = g.desc.horizon_steps, np = kModels[g.desc.model_id].np;
    rc = mas_b200_batch_create(ctx, &g.desc, B, &g.h);
    if (rc) return rc;
    std::vector<double> fx0(static_cast<size_t>(B) * n), fp, fU0;
    bool any_p = false, any_u = false;
    for (int a : g.members) {
      any_p = any_p || (model_params && model_params[a]);
      any_u = any_u || (U_init && U_init[a]);
    }
    if (any_p) fp.assign(static_cast<size_t>(B) * np, 0.0);
    if (any_u) fU0.assign(static_cast<size_t>(B) * m * T, 0.0);
    for (int sc = 0; sc < S; ++sc)
      for (int k = 0; k < G; ++k) {
        const int a = g.members[k];
        const size_t idx = static_cast<size_t>(sc) * G + k;
        std::memcpy(&fx0[idx * n], x0[a] + static_cast<size_t>(sc) * n, sizeof(double) * n);
        if (any_p)
          for (int i = 0; i < np; ++i) fp[idx * np + i] = model_params[a] ? model_params[a][static_cast<size_t>(sc) * np + i] : g.h->b->desc.params[i];
        if (any_u && U_init[a]) std::memcpy(&fU0[idx * m * T], U_init[a] + static_cast<size_t>(sc) * m * T, sizeof(double) * m * T);
      }
    rc = mas_b200_batch_set_initial_states(g.h, fx0.data());
    if (!rc) rc = mas_b200_batch_set_params(g.h, any_p ? fp.data() : nullptr);
    if (!rc) rc = mas_b200_batch_set_controls(g.h, any_u ? fU0.data() : nullptr);
    if (!rc) rc = g.h->b->initialize();
    if (!rc && keeps_old) rc = g.h->b->ensure_strategy_scratch();
    if (rc) return rc;
    if (strategy == MAS_B200_STRATEGY_TRUSTREGION) {
      std::vector<double> ones(g.h->b->ld, 1.0);
      MAS_CUDA_CHECK(cudaMemcpyAsync(g.h->b->d_radius, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      MAS_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    g.cost.assign(B, 0.0);
  }
  // joint cost of every scenario: all agents in block (= index) order, from 0.0
  std::vector<int> group_of(n_agents), slot_of(n_agents);
  for (size_t gi = 0; gi < gs.g.size(); ++gi)
    for (size_t k = 0; k < gs.g[gi].members.size(); ++k) {
      group_of[gs.g[gi].members[k]] = static_cast<int>(gi);
      slot_of[gs.g[gi].members[k]] = static_cast<int>(k);
    }
  auto download_costs = [&]() -> int {
    for (auto& g : gs.g) MAS_CUDA_CHECK(cudaMemcpyAsync(g.cost.data(), g.h->b->d_cost, g.cost.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    MAS_CUDA_CHECK(cudaStreamSynchronize(st));
    return MAS_B200_OK;
  };
  auto joint = [&](int sc) {
    double c = 0.0;
    for (int a = 0; a < n_agents; ++a) {
      const AgentGroup& g = gs.g[group_of[a]];
      c += g.cost[static_cast<size_t>(sc) * g.members.size() + slot_of[a]];
    }
    return c;
  };
  std::vector<double> base(S, 0.0);
  std::vector<int> state(S, 0);
  auto upload_state = [&]() -> int {
    for (auto& g : gs.g) MAS_CUDA_CHECK(cudaMemcpyAsync(g.h->b->d_ls_state, state.data(), S * sizeof(int), cudaMemcpyHostToDevice, st));
    MAS_CUDA_CHECK(cudaStreamSynchronize(st));
    return MAS_B200_OK;
  };
  if (strategy == MAS_B200_STRATEGY_LINESEARCH) {
    rc = download_costs();
    if (rc) return rc;
    for (int sc = 0; sc < S; ++sc) base[sc] = joint(sc);  // nash.hpp:103
  }
  std::vector<int> it_tmp;
  for (int outer = 0; outer < max_outer; ++outer) {
    for (auto& g : gs.g) {
      BatchBase* b = g.h->b;
      const size_t L = static_cast<size_t>(b->ld), nXd = L * b->nx * (b->T + 1), nUd = L * b->nu * b->T;
      if (keeps_old) {
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_U_old, b->d_U, nUd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_X_old, b->d_X, nXd * sizeof(double), cudaMemcpyDeviceToDevice, st));
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_cost_old, b->d_cost, L * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
      rc = b->solve(*params);
      if (!rc && strategy == MAS_B200_STRATEGY_TRUSTREGION) rc = b->trust_region_step();
      if (rc) return rc;
      if (trace_iterations) {
        it_tmp.resize(b->batch);
        MAS_CUDA_CHECK(cudaMemcpyAsync(it_tmp.data(), b->d_iters, b->batch * sizeof(int), cudaMemcpyDeviceToHost, st));
        MAS_CUDA_CHECK(cudaStreamSynchronize(st));
        const int G = static_cast<int>(g.members.size());
        for (int sc = 0; sc < S; ++sc)
          for (int k = 0; k < G; ++k)
            trace_iterations[(static_cast<size_t>(sc) * max_outer + outer) * n_agents + g.members[k]] = it_tmp[static_cast<size_t>(sc) * G + k];
      }
      if (strategy == MAS_B200_STRATEGY_LINESEARCH)
        MAS_CUDA_CHECK(cudaMemcpyAsync(b->d_U_cand, b->d_U, nUd * sizeof(double), cudaMemcpyDeviceToDevice, st));  // nash.hpp:123-125
    }
    if (strategy == MAS_B200_STRATEGY_LINESEARCH) {
      rc = download_costs();
      if (rc) return rc;
      bool searching = false;
      for (int sc = 0; sc < S; ++sc) {  // nash.hpp:119-121,173-176
        const double c = joint(sc);
        if (c >= base[sc]) {
          state[sc] = 1;
          searching = true;
        } else {
          base[sc] = c;
          state[sc] = 0;
        }
      }
      rc = upload_state();
      if (rc) return rc;
      for (double alpha = 0.5; searching && alpha > 1e-3; alpha *= 0.5) {  // nash.hpp:127-158
        for (auto& g : gs.g) {
          rc = g.h->b->nash_ls_trial(static_cast<int>(g.members.size()), alpha);
          if (rc) return rc;
        }
        rc = download_costs();
        if (rc) return rc;
        searching = false;
        for (int sc = 0; sc < S; ++sc) {
          if (state[sc] != 1) continue;
          const double c = joint(sc);
          if (c < base[sc]) {
            base[sc] = c;
            state[sc] = 2;
          } else {
            searching = true;
          }
        }
        rc = upload_state();
        if (rc) return rc;
      }
      for (auto& g : gs.g) {  // nash.hpp:161-171
        rc = g.h->b->nash_ls_restore(static_cast<int>(g.members.size()));
        if (rc) return rc;
      }
    }
  }
  // collect_solution (nash.hpp:23-37)
  for (auto& g : gs.g) {
    BatchBase* b = g.h->b;
    const int G = static_cast<int>(g.members.size());
    const size_t per_x = static_cast<size_t>(b->nx) * (b->T + 1), per_u = static_cast<size_t>(b->nu) * b->T;
    std::vector<double> fX(static_cast<size_t>(b->batch) * per_x), fU(static_cast<size_t>(b->batch) * per_u);
    rc = mas_b200_batch_get_solution(g.h, fX.data(), fU.data(), g.cost.data(), nullptr, nullptr);
    if (rc) return rc;
    for (int sc = 0; sc < S; ++sc)
      for (int k = 0; k < G; ++k) {
        const int a = g.members[k];
        const size_t idx = static_cast<size_t>(sc) * G + k;
        if (X && X[a]) std::memcpy(X[a] + static_cast<size_t>(sc) * per_x, &fX[idx * per_x], sizeof(double) * per_x);
        if (U && U[a]) std::memcpy(U[a] + static_cast<size_t>(sc) * per_u, &fU[idx * per_u], sizeof(double) * per_u);
        if (costs && costs[a]) costs[a][sc] = g.cost[idx];
      }
  }
  if (total_cost)
    for (int sc = 0; sc < S; ++sc) total_cost[sc] = joint(sc);
  return MAS_B200_OK;
}

int mas_b200_global_ocp_eval_mixed(mas_b200_context_t ctx, const mas_b200_ocp_desc* agent_descs, const unsigned long long* agent_ids, int n_agents,
                                   const double* X, const double* U, int time_index, double* dynamics_out, double* stage_cost_out,
                                   double* terminal_cost_out, int* dims_out, double* dt_out, double* bounds_out, int* block_agent, int* state_offsets,
                                   int* control_offsets) {
  if (!agent_descs || n_agents <= 0 || !dims_out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  if ((X || U) && !ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL (only the structure-only call works without a device)");
  for (int a = 0; a < n_agents; ++a) {
    const int rc = validate_desc(&agent_descs[a]);
    if (rc) return rc;
  }
  // compute_offsets (multi_agent_problem.hpp:37-50): blocks sorted by agent id (stable for equal ids, like the order of the input)
  std::vector<int> order(n_agents);
  for (int a = 0; a < n_agents; ++a) order[a] = a;
  if (agent_ids) std::stable_sort(order.begin(), order.end(), [&](int l, int r) { return agent_ids[l] < agent_ids[r]; });
  std::vector<int> mid(n_agents), soff(n_agents), uoff(n_agents);
  std::vector<double> prm(static_cast<size_t>(n_agents) * kMaxParams, 0.0);
  int sx = 0, su = 0;
  bool all_bounds = true;
  for (int k = 0; k < n_agents; ++k) {
    const mas_b200_ocp_desc& d = agent_descs[order[k]];
    mid[k] = d.model_id;
    soff[k] = sx;
    uoff[k] = su;
    sx += d.state_dim;
    su += d.control_dim;
    all_bounds = all_bounds && d.has_input_bounds;  // bounds only when ALL agents have both (:76-92)
    effective_params(d, &prm[static_cast<size_t>(k) * kMaxParams]);
    if (block_agent) block_agent[k] = order[k];
    if (state_offsets) state_offsets[k] = soff[k];
    if (control_offsets) control_offsets[k] = uoff[k];
  }
  dims_out[0] = sx;
  dims_out[1] = su;
  dims_out[2] = agent_descs[order[0]].horizon_steps;  // horizon and dt of the FIRST block only (:65-69)
  dims_out[3] = all_bounds ? 1 : 0;
  if (dt_out) *dt_out = agent_descs[order[0]].dt;
  if (all_bounds && bounds_out)
    for (int k = 0; k < n_agents; ++k) {
      const mas_b200_ocp_desc& d = agent_descs[order[k]];
      for (int i = 0; i < d.control_dim; ++i) {
        bounds_out[uoff[k] + i] = d.input_lower[i];
        bounds_out[su + uoff[k] + i] = d.input_upper[i];
      }
    }
  if (!X || !U) return MAS_B200_OK;  // structure only
  if (!dynamics_out || !stage_cost_out || !terminal_cost_out) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "output pointers are NULL");
  return mixed_global_eval(&ctx->c, mid.data(), soff.data(), uoff.data(), prm.data(), n_agents, sx, su, X, U, time_index, dynamics_out, stage_cost_out,
                           terminal_cost_out);
}

// ---- multi-GPU ---------------------------------------------------------------------------------------------
int mas_b200_nccl_unique_id(void* id128) {
  if (!id128) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "id128 is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
  MAS_NCCL_CHECK(g_nccl.GetUniqueId(static_cast<ncclUniqueId*>(id128)));
  return MAS_B200_OK;
}

int mas_b200_context_init_nccl(mas_b200_context_t ctx, const void* id128, int rank, int world_size) {
  if (!ctx || !id128 || world_size <= 0 || rank < 0 || rank >= world_size) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad NCCL arguments");
  int rc = load_nccl();
  if (rc) return rc;
  MAS_CUDA_CHECK(cudaSetDevice(ctx->c.device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  MAS_NCCL_CHECK(g_nccl.CommInitRank(&comm, world_size, id, rank));
  ctx->c.nccl_comm = comm;
  ctx->c.rank = rank;
  ctx->c.world = world_size;
  return MAS_B200_OK;
}

int mas_b200_context_set_blocking_sync(mas_b200_context_t ctx, int enable) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  ctx->c.blocking_sync = enable != 0;
  return MAS_B200_OK;
}

int mas_b200_context_set_agent_sharding(mas_b200_context_t ctx, int agents_sharded) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  ctx->agents_sharded = agents_sharded != 0;
  return MAS_B200_OK;
}

int mas_b200_strategy_get_joint(mas_b200_context_t ctx, double* X_all, double* U_all, double* costs_all) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  mas_b200_context::Joint& J = ctx->joint;
  if (!ctx->c.nccl_comm || !J.valid || !ctx->scratch_batch)
    return fail(MAS_B200_ERR_INVALID_ARGUMENT, "no joint trajectory set: run a Nash strategy (max_outer >= 1) on a context with a communicator first");
  BatchBase* b = ctx->scratch_batch->b;
  if (b->batch != J.batch) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "the context's batch changed since the strategy run");
  const size_t per_x = static_cast<size_t>(J.batch) * b->nx * (b->T + 1), per_u = static_cast<size_t>(J.batch) * b->nu * b->T;
  for (int rk = 0; rk < J.world; ++rk) {
    int rc = MAS_B200_OK;
    if (X_all) rc = b->download_rows(J.X + static_cast<size_t>(rk) * J.nX, X_all + rk * per_x, b->nx * (b->T + 1));
    if (!rc && U_all) rc = b->download_rows(J.U + static_cast<size_t>(rk) * J.nU, U_all + rk * per_u, b->nu * b->T);
    if (rc) return rc;
    if (costs_all)
      MAS_CUDA_CHECK(cudaMemcpyAsync(costs_all + static_cast<size_t>(rk) * J.batch, J.cost + static_cast<size_t>(rk) * J.L, J.batch * sizeof(double),
                                     cudaMemcpyDeviceToHost, ctx->c.stream));
  }
  MAS_CUDA_CHECK(cudaStreamSynchronize(ctx->c.stream));
  return MAS_B200_OK;
}

int mas_b200_strategy_get_exchange_stats(mas_b200_context_t ctx, double* collective_ms, int* rounds, long long* bytes_per_round) {
  if (!ctx) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
  const mas_b200_context::Joint& J = ctx->joint;
  if (collective_ms) *collective_ms = J.valid ? J.collective_ms : 0.0;
  if (rounds) *rounds = J.valid ? J.rounds : 0;
  if (bytes_per_round) *bytes_per_round = J.valid ? static_cast<long long>((J.nX + J.nU + J.L) * sizeof(double)) * J.world : 0;
  return MAS_B200_OK;
}

int mas_b200_synthetic_single_track_x0(unsigned long long seed, int batch, double* x0) {
  if (!x0 || batch < 0) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> dy(-2.0, 2.0), dpsi(-0.5, 0.5), dv(0.0, 2.0);
  for (int i = 0; i < batch; ++i) {
    x0[4 * i + 0] = 0.0;
    x0[4 * i + 1] = dy(rng);
    x0[4 * i + 2] = dpsi(rng);
    x0[4 * i + 3] = dv(rng);
  }
  return MAS_B200_OK;
}

int mas_b200_probe_fp64_peak(mas_b200_context_t ctx, double* tflops) {
  if (!ctx || !tflops) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  MAS_CUDA_CHECK(cudaSetDevice(ctx->c.device));
  const int blocks = ctx->c.sm_count * 8, threads = 256, iters = 20000;
  struct Scratch {  // freed on every return path
    double* d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Scratch() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
      if (d) cudaFree(d);
    }
  } sc;
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&sc.d), static_cast<size_t>(blocks) * threads * sizeof(double)));
  MAS_CUDA_CHECK(cudaEventCreate(&sc.e0));
  MAS_CUDA_CHECK(cudaEventCreate(&sc.e1));
  double* d = sc.d;
  cudaEvent_t e0 = sc.e0, e1 = sc.e1;
  dfma_probe_kernel<<<blocks, threads, 0, ctx->c.stream>>>(d, 1000);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    MAS_CUDA_CHECK(cudaEventRecord(e0, ctx->c.stream));
    dfma_probe_kernel<<<blocks, threads, 0, ctx->c.stream>>>(d, iters);
    MAS_CUDA_CHECK(cudaEventRecord(e1, ctx->c.stream));
    MAS_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    MAS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * iters * static_cast<double>(blocks) * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  *tflops = best;
  return MAS_B200_OK;
}

int mas_b200_selftest_division(mas_b200_context_t ctx, unsigned long long seed, long long pairs, long long* counts) {
  if (!ctx || !counts || pairs < 1) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "bad arguments");
  MAS_CUDA_CHECK(cudaSetDevice(ctx->c.device));
  const int threads = 256, blocks = ctx->c.sm_count * 8;
  const long long per_thread = (pairs + static_cast<long long>(threads) * blocks - 1) / (static_cast<long long>(threads) * blocks);
  if (per_thread > 1000000) return fail(MAS_B200_ERR_INVALID_ARGUMENT, "too many pairs");
  unsigned long long* d = nullptr;
  MAS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d), 5 * sizeof(unsigned long long)));
  cudaMemsetAsync(d, 0, 5 * sizeof(unsigned long long), ctx->c.stream);
  division_selftest_kernel<<<blocks, threads, 0, ctx->c.stream>>>(seed, static_cast<int>(per_thread), d);
  unsigned long long h[5] = {0, 0, 0, 0, 0};
  const cudaError_t e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->c.stream);
  const cudaError_t e2 = cudaStreamSynchronize(ctx->c.stream);
  cudaFree(d);
  MAS_CUDA_CHECK(e);
  MAS_CUDA_CHECK(e2);
  for (int i = 0; i < 5; ++i) counts[i] = static_cast<long long>(h[i]);
  return MAS_B200_OK;
}

}  // extern "C"
