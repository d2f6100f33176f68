// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).  Pinned bit for bit to a build of the reference's own sources
// (oracle/_ref, tests/test_ref_pin.py; see dense.hpp).
//
// oracle_capi.cpp: plain-C entry points over the restated reference so tests/ and bench.py's
// cpu_baseline leg can drive it through ctypes.  The batch loops are
// `#pragma omp parallel for schedule(static)` over problems / scenarios with one solver object per
// problem -- the shape of the reference's only parallel path (strategies/nash.hpp:59-64,199-202).
#include <omp.h>

#include <cstring>
#include <random>
#include <stdexcept>

#include "ref_multi_agent.hpp"
#include "ref_models.hpp"

using namespace oracle;

namespace {

enum Model { MODEL_ST_LANE = 0, MODEL_ST_CIRC = 1, MODEL_LQR = 2, MODEL_PENDULUM = 3, MODEL_ROCKET = 4, MODEL_ST_LANE_CON = 5 };

// params layout per model (all optional, NULL -> example constants):
//   ST_LANE : desired_velocity, w_lane, w_speed, w_delta, w_acc
//   ST_CIRC : track_radius, target_velocity
OCP build_ocp(int model, const double* x0, const double* params, int np, int horizon) {
  switch (model) {
    case MODEL_ST_LANE: {
      Vec s(x0, x0 + 4);
      LaneParams lp;
      if (params && np >= 5) lp = LaneParams{params[0], params[1], params[2], params[3], params[4]};
      return create_single_track_lane_following_ocp(&s, lp);
    }
    case MODEL_ST_CIRC: {
      Vec s(x0, x0 + 4);
      const double R = (params && np >= 1) ? params[0] : 20.0;
      const double v = (params && np >= 2) ? params[1] : 5.0;
      return create_single_track_circular_ocp_from_x0(s, R, v, horizon > 0 ? horizon : 10);
    }
    case MODEL_LQR: {
      Vec s(x0, x0 + 4);
      return create_linear_lqr_ocp(4, 4, 0.1, horizon > 0 ? horizon : 10, &s);
    }
    case MODEL_PENDULUM: {
      Vec s(x0, x0 + 2);
      return create_pendulum_swingup_ocp(&s);
    }
    case MODEL_ROCKET: {
      Vec s(x0, x0 + 3);
      return create_max_altitude_rocket_ocp(&s);
    }
    case MODEL_ST_LANE_CON: {  // params: desired_velocity, w_lane, w_speed, w_delta, w_acc, v_max, k_gain
      Vec s(x0, x0 + 4);
      LaneParams lp;
      double v_max = 0.8, k_gain = 0.5;
      if (params && np >= 7) {
        lp = LaneParams{params[0], params[1], params[2], params[3], params[4]};
        v_max = params[5];
        k_gain = params[6];
      }
      const int jac_mask = (params && np >= 8) ? static_cast<int>(params[7]) : 0;  // analytic constraint Jacobians (test switch)
      return create_single_track_lane_constrained_ocp(&s, lp, v_max, k_gain, jac_mask);
    }
  }
  throw std::invalid_argument("oracle: unknown model id");
}

void model_dims(int model, int horizon, int* n, int* m, int* T, double* dt) {
  switch (model) {
    case MODEL_ST_LANE: *n = 4; *m = 2; *T = 80; *dt = 0.1; break;
    case MODEL_ST_CIRC: *n = 4; *m = 2; *T = horizon > 0 ? horizon : 10; *dt = 0.5; break;
    case MODEL_LQR: *n = 4; *m = 4; *T = horizon > 0 ? horizon : 10; *dt = 0.1; break;
    case MODEL_PENDULUM: *n = 2; *m = 1; *T = 60; *dt = 0.05; break;
    case MODEL_ROCKET: *n = 3; *m = 1; *T = 50; *dt = 0.1; break;
    case MODEL_ST_LANE_CON: *n = 4; *m = 2; *T = 80; *dt = 0.1; break;
    default: throw std::invalid_argument("oracle: unknown model id");
  }
}

SolverParams make_params(int max_iterations, double tolerance, double max_ms) {
  return SolverParams{{"max_iterations", static_cast<double>(max_iterations)}, {"tolerance", tolerance}, {"max_ms", max_ms}};
}

}  // namespace

extern "C" {

int oracle_model_dims(int model, int horizon, int* n, int* m, int* T, double* dt) {
  try {
    model_dims(model, horizon, n, m, T, dt);
  } catch (...) {
    return 1;
  }
  return 0;
}

int oracle_max_threads() { return omp_get_max_threads(); }

// Config-3 initial states: x0 = (0, Y, psi, v), std::mt19937_64(seed), draws Y, psi, v per problem.
// Same generator as the product's mas_b200_synthetic_single_track_x0, here so that the CPU arm of bench.py never
// has to load the product library.
int oracle_synthetic_single_track_x0(unsigned long long seed, int batch, double* x0) {
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

// Default initial controls of the example (pendulum sinusoid, rocket constant thrust, zeros otherwise).
int oracle_default_controls(int model, int horizon, double* U /* [T][m] */) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    Vec x0(n, 0.0);
    if (model == MODEL_ROCKET) x0[2] = 1.0;
    trig_mode() = TRIG_GLIBC;  // the example computes its initial guess with libm (pendulum_swing_up.cpp:110-113)
    OCP p = build_ocp(model, x0.data(), nullptr, 0, horizon);
    std::memcpy(U, p.initial_controls.d.data(), sizeof(double) * m * T);
  } catch (...) {
    return 1;
  }
  return 0;
}

// Trajectory cost of given controls (rollout + objective), ocp.hpp:110-113,182.
int oracle_rollout_cost(int model, int batch, const double* x0, const double* params, int np, int horizon, const double* U, int trig,
                        double* X_out, double* cost_out) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    trig_mode() = trig;
    for (int b = 0; b < batch; ++b) {
      OCP p = build_ocp(model, x0 + static_cast<std::size_t>(b) * n, params ? params + static_cast<std::size_t>(b) * np : nullptr, np, horizon);
      ControlTrajectory Ub(m, T);
      std::memcpy(Ub.d.data(), U + static_cast<std::size_t>(b) * m * T, sizeof(double) * m * T);
      StateTrajectory X = integrate_horizon(p.initial_state, Ub, p.dt, p.dynamics);
      if (X_out) std::memcpy(X_out + static_cast<std::size_t>(b) * n * (T + 1), X.d.data(), sizeof(double) * n * (T + 1));
      cost_out[b] = p.objective_function(X, Ub);
    }
  } catch (...) {
    return 1;
  }
  return 0;
}

// One iLQR solve per problem (mas::solve(Solver&, OCP&), solvers/solver.hpp:28-32).
// U_inout: [batch][T][m] (= column-major m x T per problem); NULL on input side is not allowed here,
// pass the example defaults from oracle_default_controls.  stats_out (optional): per problem
// {rollouts, alpha_trials, reg_retries}.
int oracle_ilqr_solve_batch(int model, int batch, const double* x0, const double* params, int np, int horizon, double* U_inout, int max_iterations,
                            double tolerance, double max_ms, int trig, int aliased_sym, int threads, double* X_out, double* cost_out,
                            int* iters_out, int* status_out, int* stats_out) {
  int n, m, T;
  double dt;
  try {
    model_dims(model, horizon, &n, &m, &T, &dt);
  } catch (...) {
    return 1;
  }
  trig_mode() = trig;
  if (threads <= 0) threads = omp_get_max_threads();
  int err = 0;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int b = 0; b < batch; ++b) {
    try {
      OCP p = build_ocp(model, x0 + static_cast<std::size_t>(b) * n, params ? params + static_cast<std::size_t>(b) * np : nullptr, np, horizon);
      std::memcpy(p.initial_controls.d.data(), U_inout + static_cast<std::size_t>(b) * m * T, sizeof(double) * m * T);
      p.initialize_problem();
      iLQR solver;
      solver.set_params(make_params(max_iterations, tolerance, max_ms));
      solver.options.aliased_symmetrize = aliased_sym != 0;
      solver.solve(p);
      std::memcpy(U_inout + static_cast<std::size_t>(b) * m * T, p.best_controls.d.data(), sizeof(double) * m * T);
      if (X_out) std::memcpy(X_out + static_cast<std::size_t>(b) * n * (T + 1), p.best_states.d.data(), sizeof(double) * n * (T + 1));
      cost_out[b] = p.best_cost;
      if (iters_out) iters_out[b] = solver.stats.iterations;
      if (status_out) status_out[b] = solver.stats.status;
      if (stats_out) {
        stats_out[3 * b + 0] = solver.stats.rollouts;
        stats_out[3 * b + 1] = solver.stats.alpha_trials;
        stats_out[3 * b + 2] = solver.stats.reg_retries;
      }
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// The same solver object solving the same OCP n_repeat times in a row (warm start from best_controls; the
// penalty parameter and the multipliers persist across calls, ilqr.hpp:331-338,415).  Outputs after every solve.
int oracle_ilqr_solve_repeat(int model, const double* x0, const double* params, int np, int horizon, double* U_inout, int n_repeat,
                             int max_iterations, double tolerance, double penalty, int trig, double* X_out, double* cost_out, int* iters_out,
                             int* status_out) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    trig_mode() = trig;
    OCP p = build_ocp(model, x0, params, np, horizon);
    std::memcpy(p.initial_controls.d.data(), U_inout, sizeof(double) * m * T);
    p.initialize_problem();
    iLQR solver;
    SolverParams sp = make_params(max_iterations, tolerance, std::numeric_limits<double>::infinity());
    sp["penalty"] = penalty;
    solver.set_params(sp);
    for (int r = 0; r < n_repeat; ++r) {
      solver.solve(p);
      cost_out[r] = p.best_cost;
      iters_out[r] = solver.stats.iterations;
      status_out[r] = solver.stats.status;
      std::memcpy(X_out + static_cast<std::size_t>(r) * n * (T + 1), p.best_states.d.data(), sizeof(double) * n * (T + 1));
    }
    std::memcpy(U_inout, p.best_controls.d.data(), sizeof(double) * m * T);
  } catch (...) {
    return 1;
  }
  return 0;
}

// Single-problem solve with the per-iteration trace (cost after each iteration, accepted alpha index).
int oracle_ilqr_solve_trace(int model, const double* x0, const double* params, int np, int horizon, double* U_inout, int max_iterations,
                            double tolerance, int trig, int aliased_sym, double* cost_trace, int* alpha_trace, int* n_trace, int* stats3) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    trig_mode() = trig;
    OCP p = build_ocp(model, x0, params, np, horizon);
    std::memcpy(p.initial_controls.d.data(), U_inout, sizeof(double) * m * T);
    p.initialize_problem();
    iLQR solver;
    solver.set_params(make_params(max_iterations, tolerance, std::numeric_limits<double>::infinity()));
    solver.options.aliased_symmetrize = aliased_sym != 0;
    solver.solve(p);
    std::memcpy(U_inout, p.best_controls.d.data(), sizeof(double) * m * T);
    *n_trace = static_cast<int>(solver.stats.cost_trace.size());
    for (int i = 0; i < *n_trace; ++i) {
      cost_trace[i] = solver.stats.cost_trace[i];
      alpha_trace[i] = solver.stats.alpha_index[i];
    }
    stats3[0] = solver.stats.rollouts;
    stats3[1] = solver.stats.alpha_trials;
    stats3[2] = solver.stats.reg_retries;
  } catch (...) {
    return 1;
  }
  return 0;
}

// mas::solve(Strategy&, MultiAgentProblem&) (strategies/strategy.hpp:15-19) on n_scenarios
// independent scenarios of n_agents agents each.  kind: 0 centralized, 1 sequential, 2 linesearch,
// 3 trustregion.  Agents get ids 0..n_agents-1 in array order.  Arrays are [scenario][agent][...].
// trace_* (optional): [scenario][outer][agent] iterations / accepted flags / cost after the round
// (centralized: only [scenario][0][0] = iterations of the stacked solve).
int oracle_strategy_run_batch(int kind, int model, int n_scenarios, int n_agents, const double* x0, const double* params, int np, int horizon,
                              int max_outer, int max_iterations, double tolerance, double max_ms, int trig, int aliased_sym, int threads,
                              double* X_out, double* U_out, double* costs_out, double* total_cost_out, int* trace_iters, int* trace_accept,
                              double* trace_cost) {
  int n, m, T;
  double dt;
  try {
    model_dims(model, horizon, &n, &m, &T, &dt);
  } catch (...) {
    return 1;
  }
  trig_mode() = trig;
  if (threads <= 0) threads = omp_get_max_threads();
  int err = 0;
  const std::size_t per_agent_x = static_cast<std::size_t>(n) * (T + 1), per_agent_u = static_cast<std::size_t>(m) * T;
#pragma omp parallel for schedule(static) num_threads(threads)
  for (int s = 0; s < n_scenarios; ++s) {
    try {
      MultiAgentProblem problem;
      for (int a = 0; a < n_agents; ++a) {
        const std::size_t idx = static_cast<std::size_t>(s) * n_agents + a;
        auto ocp = std::make_shared<OCP>(build_ocp(model, x0 + idx * n, params ? params + idx * np : nullptr, np, horizon));
        problem.add_agent(std::make_shared<Agent>(static_cast<std::size_t>(a), ocp));
      }
      const SolverParams sp = make_params(max_iterations, tolerance, max_ms);
      OracleOptions opt;
      opt.aliased_symmetrize = aliased_sym != 0;
      StrategyTrace trace;
      Solution sol;
      if (kind == 0) {
        iLQR solver;
        solver.set_params(sp);
        solver.options = opt;
        SolveStats st;
        sol = run_centralized(solver, problem, &st);
        if (trace_iters) trace_iters[static_cast<std::size_t>(s) * max_outer * n_agents] = st.iterations;
      } else if (kind == 1) {
        sol = run_sequential(max_outer, sp, problem, opt, &trace);
      } else if (kind == 2) {
        sol = run_line_search(max_outer, sp, problem, opt, &trace);
      } else if (kind == 3) {
        sol = run_trust_region(max_outer, sp, problem, opt, &trace);
      } else {
        throw std::invalid_argument("oracle: unknown strategy kind");
      }
      for (int a = 0; a < n_agents; ++a) {
        const std::size_t idx = static_cast<std::size_t>(s) * n_agents + a;
        std::memcpy(X_out + idx * per_agent_x, sol.states[a].d.data(), sizeof(double) * per_agent_x);
        std::memcpy(U_out + idx * per_agent_u, sol.controls[a].d.data(), sizeof(double) * per_agent_u);
        costs_out[idx] = sol.costs[a];
      }
      total_cost_out[s] = sol.total_cost;
      if (kind != 0) {
        const std::size_t cnt = trace.iterations.size();
        for (std::size_t i = 0; i < cnt && i < static_cast<std::size_t>(max_outer) * n_agents; ++i) {
          const std::size_t o = static_cast<std::size_t>(s) * max_outer * n_agents + i;
          if (trace_iters) trace_iters[o] = trace.iterations[i];
          if (trace_accept) trace_accept[o] = trace.accepted[i];
          if (trace_cost) trace_cost[o] = trace.cost[i];
        }
      }
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// The same for agents of different models (MultiAgentProblem takes any mix).  Ragged arrays: per scenario the agents'
// entries are concatenated in agent order (x0: n_a; X: (T_a+1)*n_a; U: T_a*m_a); costs / iters_total are [scenario][agent].
int oracle_strategy_run_mixed(int kind, int n_scenarios, int n_agents, const int* models, const double* x0, int max_outer, int max_iterations,
                              double tolerance, int trig, double* X_out, double* U_out, double* costs_out, double* total_cost_out, int* iters_total) {
  trig_mode() = trig;
  std::vector<int> n(n_agents), m(n_agents), T(n_agents);
  std::size_t sx0 = 0, sX = 0, sU = 0;
  try {
    for (int a = 0; a < n_agents; ++a) {
      double dt;
      model_dims(models[a], 0, &n[a], &m[a], &T[a], &dt);
      if (kind == 0) T[a] = T[0];  // centralized: every agent gets its rows of the stacked solution, horizon of the first block
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
      MultiAgentProblem problem;
      std::size_t o = 0;
      for (int a = 0; a < n_agents; ++a) {
        auto ocp = std::make_shared<OCP>(build_ocp(models[a], x0 + s * sx0 + o, nullptr, 0, 0));
        o += n[a];
        problem.add_agent(std::make_shared<Agent>(static_cast<std::size_t>(a), ocp));
      }
      const SolverParams sp = make_params(max_iterations, tolerance, std::numeric_limits<double>::infinity());
      OracleOptions opt;
      StrategyTrace trace;
      Solution sol;
      int stacked_iterations = 0;
      if (kind == 0) {  // strategies/centralized.hpp:18-38 on the mixed problem
        iLQR solver;
        solver.set_params(sp);
        solver.options = opt;
        SolveStats st;
        sol = run_centralized(solver, problem, &st);
        stacked_iterations = st.iterations;
      } else if (kind == 1) sol = run_sequential(max_outer, sp, problem, opt, &trace);
      else if (kind == 2) sol = run_line_search(max_outer, sp, problem, opt, &trace);
      else if (kind == 3) sol = run_trust_region(max_outer, sp, problem, opt, &trace);
      else throw std::invalid_argument("oracle: unknown strategy kind");
      std::size_t ox = 0, ou = 0;
      for (int a = 0; a < n_agents; ++a) {
        const std::size_t px = static_cast<std::size_t>(n[a]) * (T[a] + 1), pu = static_cast<std::size_t>(m[a]) * T[a];
        std::memcpy(X_out + s * sX + ox, sol.states[a].d.data(), sizeof(double) * px);
        std::memcpy(U_out + s * sU + ou, sol.controls[a].d.data(), sizeof(double) * pu);
        ox += px;
        ou += pu;
        costs_out[static_cast<std::size_t>(s) * n_agents + a] = sol.costs[a];
        int it = 0;
        for (std::size_t r = a; r < trace.iterations.size(); r += n_agents) it += trace.iterations[r];
        if (iters_total) iters_total[static_cast<std::size_t>(s) * n_agents + a] = (kind == 0) ? (a == 0 ? stacked_iterations : 0) : it;
      }
      total_cost_out[s] = sol.total_cost;
    } catch (...) {
#pragma omp atomic write
      err = 1;
    }
  }
  return err;
}

// Stacked functions of a mixed MultiAgentProblem (build_global_ocp, multi_agent_problem.hpp:52-127): dims, horizon / dt of the
// first block, concatenated bounds when ALL agents have both, block-diagonal dynamics, costs summed in block order.
// Agents are added in REVERSE id order so that compute_offsets' sort is exercised.  bounds_out: [2][total_u] or untouched.
int oracle_global_ocp_eval_mixed(int n_agents, const int* models, const double* x0, const double* X, const double* U, double* dyn_out, double* stage_out,
                                 double* terminal_out, int* dims_out /* total_x,total_u,T,has_bounds */, double* dt_out, double* bounds_out) {
  try {
    trig_mode() = TRIG_PORTABLE;
    MultiAgentProblem problem;
    std::vector<std::shared_ptr<OCP>> ocps(n_agents);
    std::size_t o = 0;
    for (int a = 0; a < n_agents; ++a) {
      int n, m, T;
      double dt;
      model_dims(models[a], 0, &n, &m, &T, &dt);
      ocps[a] = std::make_shared<OCP>(build_ocp(models[a], x0 + o, nullptr, 0, 0));
      o += n;
    }
    for (int a = n_agents - 1; a >= 0; --a) problem.add_agent(std::make_shared<Agent>(static_cast<std::size_t>(a), ocps[a]));
    problem.compute_offsets();
    OCP g = problem.build_global_ocp();
    dims_out[0] = g.state_dim;
    dims_out[1] = g.control_dim;
    dims_out[2] = g.horizon_steps;
    dims_out[3] = (g.input_lower_bounds && g.input_upper_bounds) ? 1 : 0;
    *dt_out = g.dt;
    if (dims_out[3] && bounds_out) {
      std::memcpy(bounds_out, g.input_lower_bounds->data(), sizeof(double) * g.control_dim);
      std::memcpy(bounds_out + g.control_dim, g.input_upper_bounds->data(), sizeof(double) * g.control_dim);
    }
    Vec Xv(X, X + g.state_dim), Uv(U, U + g.control_dim);
    const Vec d = g.dynamics(Xv, Uv);
    std::memcpy(dyn_out, d.data(), sizeof(double) * d.size());
    *stage_out = g.stage_cost(Xv, Uv, 3);
    *terminal_out = g.terminal_cost(Xv);
  } catch (...) {
    return 1;
  }
  return 0;
}

// Stacked-problem evaluation used to pin build_global_ocp against tests/ocp_tests.cpp:76-154.
int oracle_global_ocp_eval(int model, int n_agents, const double* x0, const double* params, int np, int horizon, const double* X, const double* U,
                           double* dyn_out, double* stage_out, double* terminal_out, int* dims_out /* total_x,total_u,T */) {
  try {
    int n, m, T;
    double dt;
    model_dims(model, horizon, &n, &m, &T, &dt);
    MultiAgentProblem problem;
    // ids deliberately reversed so compute_offsets' sort is exercised
    for (int a = n_agents - 1; a >= 0; --a) {
      auto ocp = std::make_shared<OCP>(build_ocp(model, x0 + static_cast<std::size_t>(a) * n, params ? params + static_cast<std::size_t>(a) * np : nullptr, np, horizon));
      problem.add_agent(std::make_shared<Agent>(static_cast<std::size_t>(a), ocp));
    }
    problem.compute_offsets();
    OCP g = problem.build_global_ocp();
    dims_out[0] = g.state_dim;
    dims_out[1] = g.control_dim;
    dims_out[2] = g.horizon_steps;
    Vec Xv(X, X + g.state_dim), Uv(U, U + g.control_dim);
    const Vec d = g.dynamics(Xv, Uv);
    std::memcpy(dyn_out, d.data(), sizeof(double) * d.size());
    *stage_out = g.stage_cost(Xv, Uv, 0);
    *terminal_out = g.terminal_cost(Xv);
  } catch (...) {
    return 1;
  }
  return 0;
}

}  // extern "C"
