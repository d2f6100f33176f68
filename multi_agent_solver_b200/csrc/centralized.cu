// centralized.cu -- instantiates the stacked (centralized-strategy) solve for every registered model.
#include "centralized_host.cuh"

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

}  // namespace mas_b200
