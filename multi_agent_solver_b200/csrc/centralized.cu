// centralized.cu -- instantiates the stacked (centralized-strategy) solve for every registered model.
#include "centralized_host.cuh"
#include "stacked_mixed_host.cuh"

namespace mas_b200 {

CentralizedFn centralized_entry(int model_id) {
  switch (model_id) {
    case StLane::ID: return &run_centralized<StLane>;
    case StCirc::ID: return &run_centralized<StCirc>;
    case Lqr4::ID: return &run_centralized<Lqr4>;
    case Pendulum::ID: return &run_centralized<Pendulum>;
    case Rocket::ID: return &run_centralized<Rocket>;
    case StLaneCon::ID: return &run_centralized<StLaneCon>;  // build_global_ocp does not stack constraints
  }
  return nullptr;
}

// Centralized strategy over agents of different models (stacked_mixed.cuh).  model_ids: the agents in block order.
int centralized_mixed_entry(Context* ctx, int n_blocks, const int* model_ids, int T, double dt, int has_bounds, const double* lo, const double* hi,
                            const mas_b200_ilqr_params& prm, int S, const double* x0, const double* params, double* X, double* U, double* costs, int* ints,
                            long long* launches) {
  std::vector<MixedBlock> blocks(n_blocks);
  int ns = 0, ms = 0;
  for (int a = 0; a < n_blocks; ++a) {
    MixedBlock& b = blocks[a];
    b.model_id = model_ids[a];
    if (!mixed_model_dims(b.model_id, &b.nx, &b.nu)) {
      set_last_error("unknown model id in a mixed stacked problem");
      return MAS_B200_ERR_INVALID_ARGUMENT;
    }
    b.state_offset = ns;
    b.control_offset = ms;
    ns += b.nx;
    ms += b.nu;
    for (int i = 0; i < kMaxParams; ++i) b.params[i] = 0.0;  // the per-scenario table `params` is what the solve reads
  }
  return run_centralized_mixed(ctx, blocks, ns, ms, T, dt, has_bounds, lo, hi, prm, S, x0, params, X, U, costs, ints, launches);
}

}  // namespace mas_b200
