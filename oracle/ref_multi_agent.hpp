// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).  Pinned bit for bit to a build of the reference's own sources
// (oracle/_ref, tests/test_ref_pin.py; see dense.hpp).
//
// ref_multi_agent.hpp: CPU restatement of the reference's multi-agent layer as far as the iLQR
// path uses it:
//   agent.hpp:9-44, solution.hpp:9-15
//   multi_agent_problem.hpp:37-50 (compute_offsets), :52-127 (build_global_ocp)
//   strategies/centralized.hpp:18-38
//   strategies/nash.hpp:17-90 (sequential), :92-180 (line search), :182-248 (trust region)
// The Solver variant collapses to iLQR (the only solver on this path).
#pragma once
#include <algorithm>
#include <memory>

#include "ref_ilqr.hpp"

namespace oracle {

struct Agent {
  std::size_t id;
  std::shared_ptr<OCP> ocp;
  Agent(std::size_t id_, std::shared_ptr<OCP> o) : id(id_), ocp(std::move(o)) {}
  int state_dim() const { return ocp->state_dim; }
  int control_dim() const { return ocp->control_dim; }
  void update_initial_with_best() { ocp->update_initial_with_best(); }
};
using AgentPtr = std::shared_ptr<Agent>;

struct AgentBlockInfo {
  std::size_t agent_id;
  int state_offset, control_offset, state_dim, control_dim;
  AgentPtr agent;
};

struct Solution {
  std::vector<StateTrajectory> states;
  std::vector<ControlTrajectory> controls;
  std::vector<double> costs;
  double total_cost = 0.0;
};

inline Vec segment(const Vec& v, int off, int len) { return Vec(v.begin() + off, v.begin() + off + len); }

class MultiAgentProblem {
 public:
  std::vector<AgentPtr> agents;
  std::vector<AgentBlockInfo> blocks;
  void add_agent(const AgentPtr& a) { agents.push_back(a); }

  // multi_agent_problem.hpp:37-50
  void compute_offsets() {
    blocks.clear();
    std::vector<AgentPtr> sorted = agents;
    std::sort(sorted.begin(), sorted.end(), [](const AgentPtr& a, const AgentPtr& b) { return a->id < b->id; });
    int s_off = 0, u_off = 0;
    for (auto& a : sorted) {
      blocks.push_back({a->id, s_off, u_off, a->state_dim(), a->control_dim(), a});
      s_off += a->state_dim();
      u_off += a->control_dim();
    }
  }

  // multi_agent_problem.hpp:52-127: horizon/dt from the first block; bounds concatenated only if
  // every agent has both; block-diagonal dynamics; costs summed in block order starting from 0.0;
  // initialize_problem() then installs FD for *every* derivative (agents' analytic ones are dropped).
  OCP build_global_ocp() const {
    OCP g;
    int total_x = 0, total_u = 0;
    for (auto& b : blocks) {
      total_x += b.state_dim;
      total_u += b.control_dim;
    }
    g.state_dim = total_x;
    g.control_dim = total_u;
    if (!blocks.empty()) {
      g.horizon_steps = blocks.front().agent->ocp->horizon_steps;
      g.dt = blocks.front().agent->ocp->dt;
    }
    g.initial_state = zeros(total_x);
    for (auto& b : blocks)
      for (int i = 0; i < b.state_dim; ++i) g.initial_state[b.state_offset + i] = b.agent->ocp->initial_state[i];

    bool all_bounds = true;
    for (auto& b : blocks) all_bounds &= b.agent->ocp->input_lower_bounds.has_value() && b.agent->ocp->input_upper_bounds.has_value();
    if (all_bounds) {
      g.input_lower_bounds = zeros(total_u);
      g.input_upper_bounds = zeros(total_u);
      for (auto& b : blocks)
        for (int i = 0; i < b.control_dim; ++i) {
          (*g.input_lower_bounds)[b.control_offset + i] = (*b.agent->ocp->input_lower_bounds)[i];
          (*g.input_upper_bounds)[b.control_offset + i] = (*b.agent->ocp->input_upper_bounds)[i];
        }
    }

    g.dynamics = [bs = blocks](const State& X, const Control& U) {
      Vec out = zeros(static_cast<int>(X.size()));
      for (auto& b : bs) {
        const Vec d = b.agent->ocp->dynamics(segment(X, b.state_offset, b.state_dim), segment(U, b.control_offset, b.control_dim));
        for (int i = 0; i < b.state_dim; ++i) out[b.state_offset + i] = d[i];
      }
      return out;
    };
    g.stage_cost = [bs = blocks](const State& X, const Control& U, std::size_t t) {
      double cost = 0.0;
      for (auto& b : bs) cost += b.agent->ocp->stage_cost(segment(X, b.state_offset, b.state_dim), segment(U, b.control_offset, b.control_dim), t);
      return cost;
    };
    g.terminal_cost = [bs = blocks](const State& X) {
      double cost = 0.0;
      for (auto& b : bs) cost += b.agent->ocp->terminal_cost(segment(X, b.state_offset, b.state_dim));
      return cost;
    };
    g.initialize_problem();
    return g;
  }
};

inline Mat block_rows(const Mat& m, int row_off, int nrows) {
  Mat r(nrows, m.cols);
  for (int j = 0; j < m.cols; ++j)
    for (int i = 0; i < nrows; ++i) r(i, j) = m(row_off + i, j);
  return r;
}

// Per-round record so the device strategy kernels can be checked round by round.
struct StrategyTrace {
  std::vector<int> iterations;   // [outer * n_agents + agent]
  std::vector<int> accepted;     // trust region: 1 accepted / 0 rejected
  std::vector<double> cost;      // per-agent best_cost after the round
  std::vector<SolveStats> stats;
};

// strategies/centralized.hpp:18-38
inline Solution run_centralized(iLQR& solver, MultiAgentProblem& problem, SolveStats* stats_out = nullptr) {
  problem.compute_offsets();
  OCP global = problem.build_global_ocp();
  solver.solve(global);
  if (stats_out) *stats_out = solver.stats;
  Solution sol;
  sol.total_cost = global.best_cost;
  for (const auto& blk : problem.blocks) {
    auto& ocp = *blk.agent->ocp;
    ocp.best_states = block_rows(global.best_states, blk.state_offset, blk.state_dim);
    ocp.best_controls = block_rows(global.best_controls, blk.control_offset, blk.control_dim);
    ocp.best_cost = ocp.objective_function(ocp.best_states, ocp.best_controls);
    sol.states.push_back(ocp.best_states);
    sol.controls.push_back(ocp.best_controls);
    sol.costs.push_back(ocp.best_cost);
  }
  return sol;
}

namespace detail {

// nash.hpp:23-37: total = sum of best_cost in block order
inline Solution collect_solution(MultiAgentProblem& problem) {
  Solution sol;
  sol.total_cost = 0.0;
  for (auto& blk : problem.blocks) {
    auto& ocp = *blk.agent->ocp;
    sol.states.push_back(ocp.best_states);
    sol.controls.push_back(ocp.best_controls);
    sol.costs.push_back(ocp.best_cost);
    sol.total_cost += ocp.best_cost;
  }
  return sol;
}

// nash.hpp:39-51 (OpenMP reduction order is unspecified in the reference; block order here)
inline double total_cost(MultiAgentProblem& problem) {
  double c = 0.0;
  for (auto& blk : problem.blocks) c += blk.agent->ocp->best_cost;
  return c;
}

// nash.hpp:17-21,78-84: clones are default-constructed, then set_params
inline std::vector<iLQR> make_solvers(std::size_t n, const SolverParams& params, const OracleOptions& opt) {
  std::vector<iLQR> solvers(n);
  for (auto& s : solvers) {
    s.set_params(params);
    s.options = opt;
  }
  return solvers;
}

// nash.hpp:53-72
inline void sequential_solve(std::vector<iLQR>& solvers, MultiAgentProblem& problem, StrategyTrace* trace) {
  const int n = static_cast<int>(problem.blocks.size());
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) solvers[i].solve(*problem.blocks[i].agent->ocp);
  for (int i = 0; i < n; ++i) problem.blocks[i].agent->update_initial_with_best();
  if (trace)
    for (int i = 0; i < n; ++i) {
      trace->iterations.push_back(solvers[i].stats.iterations);
      trace->accepted.push_back(1);
      trace->cost.push_back(problem.blocks[i].agent->ocp->best_cost);
      trace->stats.push_back(solvers[i].stats);
    }
}

}  // namespace detail

// nash.hpp:74-90: exactly max_outer Jacobi rounds, no stop test
inline Solution run_sequential(int max_outer, const SolverParams& params, MultiAgentProblem& problem, const OracleOptions& opt = {},
                               StrategyTrace* trace = nullptr) {
  problem.compute_offsets();
  auto solvers = detail::make_solvers(problem.blocks.size(), params, opt);
  for (int outer = 0; outer < max_outer; ++outer) detail::sequential_solve(solvers, problem, trace);
  return detail::collect_solution(problem);
}

// nash.hpp:92-180
inline Solution run_line_search(int max_outer, const SolverParams& params, MultiAgentProblem& problem, const OracleOptions& opt = {},
                                StrategyTrace* trace = nullptr) {
  problem.compute_offsets();
  auto solvers = detail::make_solvers(problem.blocks.size(), params, opt);
  double base_cost = detail::total_cost(problem);
  for (int outer = 0; outer < max_outer; ++outer) {
    const int n = static_cast<int>(problem.blocks.size());
    std::vector<ControlTrajectory> old_controls(n);
    std::vector<StateTrajectory> old_states(n);
    for (int i = 0; i < n; ++i) {
      old_controls[i] = problem.blocks[i].agent->ocp->best_controls;
      old_states[i] = problem.blocks[i].agent->ocp->best_states;
    }
    detail::sequential_solve(solvers, problem, trace);
    const double new_cost = detail::total_cost(problem);
    if (new_cost >= base_cost) {
      std::vector<ControlTrajectory> cand_controls(n);
      for (int i = 0; i < n; ++i) cand_controls[i] = problem.blocks[i].agent->ocp->best_controls;
      double alpha = 0.5;
      bool accepted = false;
      while (alpha > 1e-3 && !accepted) {
        std::vector<ControlTrajectory> trial_controls(n);
        std::vector<StateTrajectory> trial_states(n);
        double trial_cost = 0.0;
        for (int i = 0; i < n; ++i) {
          auto& ocp = *problem.blocks[i].agent->ocp;
          trial_controls[i] = add(old_controls[i], scale(alpha, sub(cand_controls[i], old_controls[i])));
          trial_states[i] = integrate_horizon(ocp.initial_state, trial_controls[i], ocp.dt, ocp.dynamics);
          trial_cost += ocp.objective_function(trial_states[i], trial_controls[i]);
        }
        if (trial_cost < base_cost) {
          for (int i = 0; i < n; ++i) {
            auto& ocp = *problem.blocks[i].agent->ocp;
            ocp.best_controls = trial_controls[i];
            ocp.best_states = trial_states[i];
            ocp.best_cost = ocp.objective_function(trial_states[i], trial_controls[i]);
            ocp.update_initial_with_best();
          }
          base_cost = trial_cost;
          accepted = true;
        } else {
          alpha *= 0.5;
        }
      }
      if (!accepted) {
        for (int i = 0; i < n; ++i) {
          auto& ocp = *problem.blocks[i].agent->ocp;
          ocp.best_controls = old_controls[i];
          ocp.best_states = old_states[i];
          ocp.best_cost = ocp.objective_function(old_states[i], old_controls[i]);
          ocp.update_initial_with_best();
        }
      }
    } else {
      base_cost = new_cost;
    }
  }
  return detail::collect_solution(problem);
}

// nash.hpp:182-248
inline Solution run_trust_region(int max_outer, const SolverParams& params, MultiAgentProblem& problem, const OracleOptions& opt = {},
                                 StrategyTrace* trace = nullptr) {
  problem.compute_offsets();
  auto solvers = detail::make_solvers(problem.blocks.size(), params, opt);
  std::vector<double> radii(problem.blocks.size(), 1.0);
  for (int outer = 0; outer < max_outer; ++outer) {
    const int n = static_cast<int>(problem.blocks.size());
    std::vector<int> acc(n, 0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      auto& ocp = *problem.blocks[i].agent->ocp;
      const ControlTrajectory old_u = ocp.best_controls;
      const StateTrajectory old_x = ocp.best_states;
      const double old_cost = ocp.best_cost;

      solvers[i].solve(ocp);

      ControlTrajectory cand_u = ocp.best_controls;
      StateTrajectory cand_x = ocp.best_states;
      double cand_cost = ocp.best_cost;

      const ControlTrajectory delta = sub(cand_u, old_u);
      const double norm = frobenius_norm(delta);
      if (norm > radii[i]) {
        const double s = radii[i] / norm;
        cand_u = add(old_u, scale(s, delta));
        cand_x = integrate_horizon(ocp.initial_state, cand_u, ocp.dt, ocp.dynamics);
        cand_cost = ocp.objective_function(cand_x, cand_u);
      }
      if (cand_cost < old_cost) {
        ocp.best_controls = cand_u;
        ocp.best_states = cand_x;
        ocp.best_cost = cand_cost;
        ocp.update_initial_with_best();
        radii[i] *= 1.5;
        acc[i] = 1;
      } else {
        ocp.best_controls = old_u;
        ocp.best_states = old_x;
        ocp.best_cost = old_cost;
        ocp.update_initial_with_best();
        radii[i] *= 0.5;
      }
    }
    if (trace)
      for (int i = 0; i < n; ++i) {
        trace->iterations.push_back(solvers[i].stats.iterations);
        trace->accepted.push_back(acc[i]);
        trace->cost.push_back(problem.blocks[i].agent->ocp->best_cost);
        trace->stats.push_back(solvers[i].stats);
      }
  }
  return detail::collect_solution(problem);
}

}  // namespace oracle
