// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the product.
//
// ref_capi.cpp: plain-C entry points over a build of the reference's OWN sources
// (/root/reference/include/multi_agent_solver/** and examples/*.cpp, unmodified, compiled against
// oracle/eigen_shim) so that tests/ can assert  oracle == reference  and bench.py can time the
// reference's CPU implementation (cpu_baseline.kind = "reference").
//
// The reference's solve() returns nothing, so the counters the oracle defines (SURVEY 8a:
// iterations = loop bodies that reached the backward pass; status; alpha trials; regularisation
// retries) are taken here WITHOUT editing the reference: the OCP's public callbacks are wrapped --
// terminal_cost_gradient is called exactly once per iteration that reaches the backward pass
// (ilqr.hpp:92-93), objective_function once per line-search candidate and once after it
// (ilqr.hpp:220,233) -- and the shim counts LLT::compute calls (ilqr.hpp:175-182).
#include <omp.h>

#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <vector>

#include "multi_agent_solver/multi_agent_solver.hpp"
#include "ref_examples.hpp"

extern "C" void ref_set_trig_mode(int mode);
extern "C" void ref_set_thread_trig_override(int mode);

namespace {

enum Model { MODEL_ST_LANE = 0, MODEL_ST_CIRC = 1, MODEL_LQR = 2, MODEL_PENDULUM = 3, MODEL_ROCKET = 4, MODEL_ST_LANE_CON = 5 };
enum Status { STATUS_CONVERGED = 0, STATUS_MAX_ITER = 1, STATUS_TIME_LIMIT = 2 };

void model_dims(int model, int horizon, int* n, int* m, int* T, double* dt) {
  switch (model) {
    case MODEL_ST_LANE: case MODEL_ST_LANE_CON: *n = 4; *m = 2; *T = 80; *dt = 0.1; break;
    case MODEL_ST_CIRC: *n = 4; *m = 2; *T = horizon > 0 ? horizon : 10; *dt = 0.5; break;
    case MODEL_LQR: *n = 4; *m = 4; *T = horizon > 0 ? horizon : 10; *dt = 0.1; break;
    case MODEL_PENDULUM: *n = 2; *m = 1; *T = 60; *dt = 0.05; break;
    case MODEL_ROCKET: *n = 3; *m = 1; *T = 50; *dt = 0.1; break;
    default: throw std::invalid_argument("ref: unknown model id");
  }
}

// The example's builder, then the caller's initial state (and optionally controls) through the
// OCP's public members and a second initialize_problem() (ocp.hpp:102-183), as a user would.
// params: ST_CIRC {track_radius, target_velocity}; ST_LANE_CON {v_max, k_gain}; others: none (the
// example constants are compiled into the reference's lambdas).
mas::OCP build_ocp(int model, const double* x0, const double* params, int np, int horizon, const double* U /* [T][m] or null */) {
  mas::OCP p;
  ref_set_thread_trig_override(0);  // the example's own set-up code runs with libm, whatever mode the solve is compared in
  struct Restore {
    ~Restore() { ref_set_thread_trig_override(-1); }
  } restore_after_builder;
  switch (model) {
    case MODEL_ST_LANE: p = create_single_track_lane_following_ocp(); break;
    case MODEL_ST_CIRC:
      p = create_single_track_circular_ocp(0.0, (params && np >= 1) ? params[0] : 20.0, (params && np >= 2) ? params[1] : 5.0,
                                           horizon > 0 ? horizon : 10);
      break;
    case MODEL_LQR: p = create_linear_lqr_ocp(4, 4, 0.1, horizon > 0 ? horizon : 10); break;
    case MODEL_PENDULUM: p = create_pendulum_swingup_ocp(); break;
    case MODEL_ROCKET: p = mas::create_max_altitude_rocket_ocp(); break;
    case MODEL_ST_LANE_CON: {
      // Not a reference example (no example sets constraints): the reference's lane-following OCP plus
      // the path constraints of the repo's test model, given through the OCP's public constraint
      // members so that the reference's augmented-Lagrangian branch (ilqr.hpp:121-170,236-260,380-407)
      // runs.  eq: a - k (v_des - v) = 0 ; ineq: v - v_max <= 0.
      p = create_single_track_lane_following_ocp();
      const double v_max = (params && np >= 7) ? params[5] : 0.8, k_gain = (params && np >= 7) ? params[6] : 0.5;
      const double v_des = 1.0;
      p.equality_constraints = [=](const mas::State& x, const mas::Control& u) {
        mas::ConstraintViolations c(1);
        c(0) = u(1) - k_gain * (v_des - x(3));
        return c;
      };
      p.inequality_constraints = [=](const mas::State& x, const mas::Control&) {
        mas::ConstraintViolations c(1);
        c(0) = x(3) - v_max;
        return c;
      };
      // params[7] (test switch): which constraint Jacobians are given analytically (ocp.hpp:65-68): 1 eq/state, 2 eq/control,
      // 4 ineq/state, 8 ineq/control; the others get the finite-difference defaults from initialize_problem (ocp.hpp:137-171)
      const int jac_mask = (params && np >= 8) ? static_cast<int>(params[7]) : 0;
      auto row = [](std::initializer_list<double> v) {
        mas::ConstraintsJacobian J = mas::ConstraintsJacobian::Zero(1, static_cast<int>(v.size()));
        int c = 0;
        for (double e : v) J(0, c++) = e;
        return J;
      };
      if (jac_mask & 1) p.equality_constraints_state_jacobian = [=](const mas::State&, const mas::Control&) { return row({0.0, 0.0, 0.0, k_gain}); };
      if (jac_mask & 2) p.equality_constraints_control_jacobian = [=](const mas::State&, const mas::Control&) { return row({0.0, 1.0}); };
      if (jac_mask & 4) p.inequality_constraints_state_jacobian = [=](const mas::State&, const mas::Control&) { return row({0.0, 0.0, 0.0, 1.0}); };
      if (jac_mask & 8) p.inequality_constraints_control_jacobian = [=](const mas::State&, const mas::Control&) { return row({0.0, 0.0}); };
      break;
    }
    default: throw std::invalid_argument("ref: unknown model id");
  }
  ref_set_thread_trig_override(-1);
  for (int i = 0; i < p.state_dim; ++i) p.initial_state(i) = x0[i];
  if (U) {
    p.initial_controls.resize(p.control_dim, p.horizon_steps);
    std::memcpy(p.initial_controls.data(), U, sizeof(double) * p.control_dim * p.horizon_steps);
  }
  p.initialize_problem();
  return p;
}

struct Probe {
  std::vector<double> objective_values;   // every objective_function result, in call order
  std::vector<std::size_t> iteration_at;  // objective_values.size() when an iteration's backward pass began
};

std::shared_ptr<Probe> instrument(mas::OCP& p) {
  auto probe = std::make_shared<Probe>();
  auto tg = p.terminal_cost_gradient;
  p.terminal_cost_gradient = [tg, probe](const mas::TerminalCostFunction& f, const mas::State& x) {
    probe->iteration_at.push_back(probe->objective_values.size());
    return tg(f, x);
  };
  auto obj = p.objective_function;
  p.objective_function = [obj, probe](const mas::StateTrajectory& X, const mas::ControlTrajectory& U) {
    const double v = obj(X, U);
    probe->objective_values.push_back(v);
    return v;
  };
  return probe;
}

mas::SolverParams make_params(int max_iterations, double tolerance, double max_ms) {
  return mas::SolverParams{{"max_iterations", static_cast<double>(max_iterations)}, {"tolerance", tolerance}, {"max_ms", max_ms}};
}

// iterations / trials / status from the probe of ONE unconstrained solve (merit == objective):
// objective calls are  cost, merit0, then per iteration  trial_1 .. trial_k, cost.
void decode_probe(const Probe& pr, std::size_t first_obj, std::size_t first_iter, int max_iterations, double tolerance, bool constrained,
                  int* iterations, int* alpha_trials, int* status) {
  const int iters = static_cast<int>(pr.iteration_at.size() - first_iter);
  const int n_obj = static_cast<int>(pr.objective_values.size() - first_obj);
  *iterations = iters;
  *alpha_trials = constrained ? -1 : n_obj - 2 - iters;
  if (constrained) {  // merit != objective and compute_merit also calls the objective: not decodable here
    *status = -1;
    return;
  }
  int st = iters >= max_iterations ? STATUS_MAX_ITER : STATUS_TIME_LIMIT;
  if (iters > 0) {
    double merit = pr.objective_values[first_obj + 1];
    bool last_converged = false;
    for (int it = 0; it < iters; ++it) {
      const std::size_t b = pr.iteration_at[first_iter + it];
      const std::size_t e = (it + 1 < iters) ? pr.iteration_at[first_iter + it + 1] : pr.objective_values.size();
      // [b, e-1) are the candidates, e-1 is the recomputed cost; only the last candidate can have been accepted
      double best = merit;
      if (e - b >= 2) {
        const double last_trial = pr.objective_values[e - 2];
        if (last_trial < merit) best = last_trial;
      }
      last_converged = (merit - best) < tolerance;
      merit = best;
    }
    if (last_converged) st = STATUS_CONVERGED;
  }
  *status = st;
}

}  // namespace

extern "C" {

int ref_max_threads() { return omp_get_max_threads(); }

// Config-3 initial states: x0 = (0, Y, psi, v), std::mt19937_64(seed), draws Y, psi, v per problem.
// Same generator as the product's mas_b200_synthetic_single_track_x0, here so that the CPU arm of bench.py never
// has to load the product library.
int ref_synthetic_single_track_x0(unsigned long long seed, int batch, double* x0) {
  std::mt19937_64 rng(seed);
  std::uniform_real_distribution<double> dy(-2.0, 2.0), dpsi(-0.5, 0.5), dv(0.0, 2.0);
  for (int i = 0; i < batch; ++i) {
    x0[4 * i + 0] = 0.0;
    x0[4 * i + 1] = dy(rng);
    x0[4 * i + 2] = dpsi(rng);
    x0[4 * i + 3] = dv(rng);
  }
  return 0;
}

int ref_model_dims(int model, int horizon, int* n, int* m, int* T, double* dt) {
  try {
    model_dims(model, horizon, n, m, T, dt);
  } catch (...) {
    return 1;
  }
  return 0;
}

// Example default controls (pendulum sinusoid, rocket constant thrust, zeros otherwise).
int ref_default_controls(int model, int horizon, double* U) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    std::vector<double> x0(n, 0.0);
    if (model == MODEL_ROCKET) x0[2] = 1.0;
    ref_set_trig_mode(0);  // the example computes its initial guess with libm (pendulum_swing_up.cpp:110-113)
    mas::OCP p = build_ocp(model, x0.data(), nullptr, 0, horizon, nullptr);
    std::memcpy(U, p.initial_controls.data(), sizeof(double) * m * T);
  } catch (...) {
    return 1;
  }
  return 0;
}

// mas::solve(Solver&, OCP&) (solvers/solver.hpp:28-32) with Solver = iLQR, one fresh solver per problem,
// `#pragma omp parallel for schedule(static)` over problems like the reference's only parallel path
// (strategies/nash.hpp:59-64).  instrument = 0: no callback wrapping at all (used for timing).
int ref_ilqr_solve_batch(int model, int batch, const double* x0, const double* params, int np, int horizon, double* U_inout, int max_iterations,
                         double tolerance, double max_ms, int trig, int threads, int do_instrument, double* X_out, double* cost_out,
                         int* iters_out, int* status_out, int* stats_out) {
  int n, m, T;
  double dt;
  try {
    model_dims(model, horizon, &n, &m, &T, &dt);
  } catch (...) {
    return 1;
  }
  ref_set_trig_mode(trig);
  if (threads <= 0) threads = omp_get_max_threads();
  int err = 0;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int b = 0; b < batch; ++b) {
    try {
      mas::OCP p = build_ocp(model, x0 + static_cast<std::size_t>(b) * n, params ? params + static_cast<std::size_t>(b) * np : nullptr, np,
                             horizon, U_inout + static_cast<std::size_t>(b) * m * T);
      std::shared_ptr<Probe> probe;
      if (do_instrument) probe = instrument(p);
      mas::Solver solver{std::in_place_type<mas::iLQR>};
      mas::set_params(solver, make_params(max_iterations, tolerance, max_ms));
      const long llt0 = Eigen::shim::llt_compute_calls();
      mas::solve(solver, p);
      const long llt = Eigen::shim::llt_compute_calls() - llt0;
      std::memcpy(U_inout + static_cast<std::size_t>(b) * m * T, p.best_controls.data(), sizeof(double) * m * T);
      if (X_out) std::memcpy(X_out + static_cast<std::size_t>(b) * n * (T + 1), p.best_states.data(), sizeof(double) * n * (T + 1));
      cost_out[b] = p.best_cost;
      if (probe) {
        int it, tr, st;
        decode_probe(*probe, 0, 0, max_iterations, tolerance, model == MODEL_ST_LANE_CON, &it, &tr, &st);
        if (iters_out) iters_out[b] = it;
        if (status_out) status_out[b] = st;
        if (stats_out) {
          stats_out[3 * b + 0] = tr < 0 ? -1 : 1 + tr;                    // rollouts incl. the prologue one
          stats_out[3 * b + 1] = tr;                                      // line-search candidates
          stats_out[3 * b + 2] = static_cast<int>(llt - static_cast<long>(T) * it);  // Q_uu + reg*I retries
        }
      }
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// One solver object, n_repeat solve() calls on the same OCP (multipliers / penalty persist, ilqr.hpp:331-338).
int ref_ilqr_solve_repeat(int model, const double* x0, const double* params, int np, int horizon, double* U_inout, int n_repeat,
                          int max_iterations, double tolerance, double penalty, int trig, double* X_out, double* cost_out, int* iters_out) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    ref_set_trig_mode(trig);
    mas::OCP p = build_ocp(model, x0, params, np, horizon, U_inout);
    auto probe = instrument(p);
    mas::Solver solver{std::in_place_type<mas::iLQR>};
    mas::SolverParams sp = make_params(max_iterations, tolerance, std::numeric_limits<double>::infinity());
    sp["penalty"] = penalty;
    mas::set_params(solver, sp);
    for (int r = 0; r < n_repeat; ++r) {
      const std::size_t it0 = probe->iteration_at.size();
      mas::solve(solver, p);
      cost_out[r] = p.best_cost;
      iters_out[r] = static_cast<int>(probe->iteration_at.size() - it0);
      std::memcpy(X_out + static_cast<std::size_t>(r) * n * (T + 1), p.best_states.data(), sizeof(double) * n * (T + 1));
    }
    std::memcpy(U_inout, p.best_controls.data(), sizeof(double) * m * T);
  } catch (...) {
    return 1;
  }
  return 0;
}

// mas::solve(Strategy&, MultiAgentProblem&) (strategies/strategy.hpp:15-19) through the reference's own
// registry (examples/example_utils.hpp:94-110 is a header of the examples; here the strategy objects are
// built directly with the same constructor arguments).  kind: 0 centralized, 1 sequential, 2 linesearch,
// 3 trustregion.  Arrays are [scenario][agent][...]; iters_total_out[scenario][agent] = iLQR iterations
// summed over the outer rounds (centralized: [scenario][0] = iterations of the stacked solve, measured on a
// second, instrumented build_global_ocp() + solve, only when iters_total_out is given).
int ref_strategy_run_batch(int kind, int model, int n_scenarios, int n_agents, const double* x0, const double* params, int np, int horizon,
                           int max_outer, int max_iterations, double tolerance, double max_ms, int trig, int threads, double* X_out,
                           double* U_out, double* costs_out, double* total_cost_out, int* iters_total_out) {
  int n, m, T;
  double dt;
  try {
    model_dims(model, horizon, &n, &m, &T, &dt);
  } catch (...) {
    return 1;
  }
  ref_set_trig_mode(trig);
  if (threads <= 0) threads = omp_get_max_threads();
  int err = 0;
  const std::size_t per_agent_x = static_cast<std::size_t>(n) * (T + 1), per_agent_u = static_cast<std::size_t>(m) * T;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int s = 0; s < n_scenarios; ++s) {
    try {
      // the reference's own `omp parallel for` loops over agents (nash.hpp:45,59,134,199) run inside this
      // scenario on one thread: sequential block order, the order a one-thread reference run has
      omp_set_num_threads(1);
      auto make_problem = [&](std::vector<std::shared_ptr<Probe>>* probes) {
        mas::MultiAgentProblem problem;
        for (int a = 0; a < n_agents; ++a) {
          const std::size_t idx = static_cast<std::size_t>(s) * n_agents + a;
          auto ocp = std::make_shared<mas::OCP>(build_ocp(model, x0 + idx * n, params ? params + idx * np : nullptr, np, horizon, nullptr));
          if (probes) probes->push_back(instrument(*ocp));
          problem.add_agent(std::make_shared<mas::Agent>(static_cast<std::size_t>(a), ocp));
        }
        return problem;
      };
      std::vector<std::shared_ptr<Probe>> probes;
      mas::MultiAgentProblem problem = make_problem(kind == 0 ? nullptr : &probes);
      const mas::SolverParams sp = make_params(max_iterations, tolerance, max_ms);
      mas::Solver solver{std::in_place_type<mas::iLQR>};
      mas::Strategy strategy = [&]() -> mas::Strategy {
        switch (kind) {
          case 0: mas::set_params(solver, sp); return mas::CentralizedStrategy{std::move(solver)};
          case 1: return mas::SequentialNashStrategy{max_outer, std::move(solver), sp};
          case 2: return mas::LineSearchNashStrategy{max_outer, std::move(solver), sp};
          case 3: return mas::TrustRegionNashStrategy{max_outer, std::move(solver), sp};
        }
        throw std::invalid_argument("ref: unknown strategy kind");
      }();
      const mas::Solution sol = mas::solve(strategy, problem);
      for (int a = 0; a < n_agents; ++a) {
        const std::size_t idx = static_cast<std::size_t>(s) * n_agents + a;
        std::memcpy(X_out + idx * per_agent_x, sol.states[a].data(), sizeof(double) * per_agent_x);
        std::memcpy(U_out + idx * per_agent_u, sol.controls[a].data(), sizeof(double) * per_agent_u);
        costs_out[idx] = sol.costs[a];
        if (iters_total_out && kind != 0) iters_total_out[idx] = static_cast<int>(probes[a]->iteration_at.size());
      }
      total_cost_out[s] = sol.total_cost;
      if (iters_total_out && kind == 0) {
        mas::MultiAgentProblem again = make_problem(nullptr);
        again.compute_offsets();
        mas::OCP global = again.build_global_ocp();
        auto probe = instrument(global);
        mas::Solver s2{std::in_place_type<mas::iLQR>};
        mas::set_params(s2, sp);
        mas::solve(s2, global);
        iters_total_out[static_cast<std::size_t>(s) * n_agents] = static_cast<int>(probe->iteration_at.size());
      }
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// Agents of different models in one MultiAgentProblem (ragged arrays as in oracle_strategy_run_mixed).
int ref_strategy_run_mixed(int kind, int n_scenarios, int n_agents, const int* models, const double* x0, int max_outer, int max_iterations,
                           double tolerance, int trig, double* X_out, double* U_out, double* costs_out, double* total_cost_out, int* iters_total) {
  ref_set_trig_mode(trig);
  std::vector<int> n(n_agents), m(n_agents), T(n_agents);
  std::size_t sx0 = 0, sX = 0, sU = 0;
  try {
    for (int a = 0; a < n_agents; ++a) {
      double dt;
      model_dims(models[a], 0, &n[a], &m[a], &T[a], &dt);
      if (kind == 0) T[a] = T[0];  // centralized: rows of the stacked solution, horizon of the first block
      sx0 += n[a];
      sX += static_cast<std::size_t>(n[a]) * (T[a] + 1);
      sU += static_cast<std::size_t>(m[a]) * T[a];
    }
  } catch (...) {
    return 1;
  }
  int err = 0;
#pragma omp parallel for schedule(static)
  for (int s = 0; s < n_scenarios; ++s) {
    try {
      omp_set_num_threads(1);
      std::vector<std::shared_ptr<Probe>> probes;
      auto make_problem = [&](bool with_probes) {
        mas::MultiAgentProblem pr;
        std::size_t o = 0;
        for (int a = 0; a < n_agents; ++a) {
          auto ocp = std::make_shared<mas::OCP>(build_ocp(models[a], x0 + s * sx0 + o, nullptr, 0, 0, nullptr));
          o += n[a];
          if (with_probes) probes.push_back(instrument(*ocp));
          pr.add_agent(std::make_shared<mas::Agent>(static_cast<std::size_t>(a), ocp));
        }
        return pr;
      };
      mas::MultiAgentProblem problem = make_problem(kind != 0);
      const mas::SolverParams sp = make_params(max_iterations, tolerance, std::numeric_limits<double>::infinity());
      mas::Solver solver{std::in_place_type<mas::iLQR>};
      mas::Strategy strategy = [&]() -> mas::Strategy {
        switch (kind) {
          case 0: mas::set_params(solver, sp); return mas::CentralizedStrategy{std::move(solver)};
          case 1: return mas::SequentialNashStrategy{max_outer, std::move(solver), sp};
          case 2: return mas::LineSearchNashStrategy{max_outer, std::move(solver), sp};
          case 3: return mas::TrustRegionNashStrategy{max_outer, std::move(solver), sp};
        }
        throw std::invalid_argument("ref: unknown strategy kind");
      }();
      const mas::Solution sol = mas::solve(strategy, problem);
      std::size_t ox = 0, ou = 0;
      for (int a = 0; a < n_agents; ++a) {
        const std::size_t px = static_cast<std::size_t>(n[a]) * (T[a] + 1), pu = static_cast<std::size_t>(m[a]) * T[a];
        std::memcpy(X_out + s * sX + ox, sol.states[a].data(), sizeof(double) * px);
        std::memcpy(U_out + s * sU + ou, sol.controls[a].data(), sizeof(double) * pu);
        ox += px;
        ou += pu;
        costs_out[static_cast<std::size_t>(s) * n_agents + a] = sol.costs[a];
        if (iters_total) iters_total[static_cast<std::size_t>(s) * n_agents + a] = kind == 0 ? 0 : static_cast<int>(probes[a]->iteration_at.size());
      }
      total_cost_out[s] = sol.total_cost;
      if (iters_total && kind == 0) {  // iterations of the stacked solve: a second, instrumented build_global_ocp() + solve
        mas::MultiAgentProblem again = make_problem(false);
        again.compute_offsets();
        mas::OCP global = again.build_global_ocp();
        auto probe = instrument(global);
        mas::Solver s2{std::in_place_type<mas::iLQR>};
        mas::set_params(s2, sp);
        mas::solve(s2, global);
        iters_total[static_cast<std::size_t>(s) * n_agents] = static_cast<int>(probe->iteration_at.size());
      }
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// compute_offsets + build_global_ocp of a mixed problem (agents added in reverse id order), evaluated at (X, U); portable trig.
int ref_global_ocp_eval_mixed(int n_agents, const int* models, const double* x0, const double* X, const double* U, double* dyn_out, double* stage_out,
                              double* terminal_out, int* dims_out, double* dt_out, double* bounds_out) {
  try {
    ref_set_trig_mode(1);
    mas::MultiAgentProblem problem;
    std::vector<std::shared_ptr<mas::OCP>> ocps(n_agents);
    std::size_t o = 0;
    for (int a = 0; a < n_agents; ++a) {
      int n, m, T;
      double dt;
      model_dims(models[a], 0, &n, &m, &T, &dt);
      ocps[a] = std::make_shared<mas::OCP>(build_ocp(models[a], x0 + o, nullptr, 0, 0, nullptr));
      o += n;
    }
    for (int a = n_agents - 1; a >= 0; --a) problem.add_agent(std::make_shared<mas::Agent>(static_cast<std::size_t>(a), ocps[a]));
    problem.compute_offsets();
    mas::OCP g = problem.build_global_ocp();
    dims_out[0] = g.state_dim;
    dims_out[1] = g.control_dim;
    dims_out[2] = g.horizon_steps;
    dims_out[3] = (g.input_lower_bounds && g.input_upper_bounds) ? 1 : 0;
    *dt_out = g.dt;
    if (dims_out[3] && bounds_out) {
      std::memcpy(bounds_out, g.input_lower_bounds->data(), sizeof(double) * g.control_dim);
      std::memcpy(bounds_out + g.control_dim, g.input_upper_bounds->data(), sizeof(double) * g.control_dim);
    }
    mas::State Xv(g.state_dim);
    mas::Control Uv(g.control_dim);
    for (int i = 0; i < g.state_dim; ++i) Xv(i) = X[i];
    for (int i = 0; i < g.control_dim; ++i) Uv(i) = U[i];
    const mas::StateDerivative d = g.dynamics(Xv, Uv);
    std::memcpy(dyn_out, d.data(), sizeof(double) * d.size());
    *stage_out = g.stage_cost(Xv, Uv, 3);
    *terminal_out = g.terminal_cost(Xv);
  } catch (...) {
    return 1;
  }
  return 0;
}

}  // extern "C"
