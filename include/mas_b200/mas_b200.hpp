// mas_b200.hpp -- host-side C++ facade over the C ABI (include/mas_b200.h).
//
// Mirrors the reference's interface for the iLQR path -- same names, argument meaning and error
// behaviour -- so code written against markomiz/multi_agent_solver ports by swapping the namespace:
//
//   reference (include/multi_agent_solver/...)           here (namespace mas_b200)
//   ocp.hpp:30-237            struct OCP                 struct OCP        (callbacks -> model_id + params + deriv_mask)
//   solvers/ilqr.hpp:23-55    class iLQR                 class iLQR        (set_params / solve)
//   solvers/solver.hpp:17-45  Solver, solve, set_params  same, variant over the device solvers
//   agent.hpp, solution.hpp   Agent, Solution            same
//   multi_agent_problem.hpp   MultiAgentProblem          add_agent / compute_offsets / blocks
//   strategies/*.hpp          four strategy functors     same constructors, Solution operator()(MultiAgentProblem&)
//   strategies/strategy.hpp   Strategy, solve            same
//   examples/example_utils.hpp:32-110  name registries   canonical_*_name, make_solver, make_strategy
//
// Header-only; link against libmas_b200.so.  Define MAS_B200_NAMESPACE_ALIAS_MAS to get `namespace mas`.
// Everything runs on the GPU through the C ABI: there is no host implementation of the algorithm here.
#pragma once

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <variant>
#include <vector>

#include "mas_b200.h"

namespace mas_b200 {

using State = std::vector<double>;
using Control = std::vector<double>;
using SolverParams = std::unordered_map<std::string, double>;  // types.hpp:57

// Column-major dense matrix with the handful of Eigen::MatrixXd accessors the reference's callers use.
class Matrix {
 public:
  Matrix() = default;
  Matrix(int rows, int cols, double value = 0.0) : rows_(rows), cols_(cols), d_(static_cast<std::size_t>(rows) * cols, value) {}
  static Matrix Zero(int rows, int cols) { return Matrix(rows, cols, 0.0); }
  static Matrix Constant(int rows, int cols, double v) { return Matrix(rows, cols, v); }
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  std::size_t size() const { return d_.size(); }
  double& operator()(int i, int j) { return d_[static_cast<std::size_t>(j) * rows_ + i]; }
  double operator()(int i, int j) const { return d_[static_cast<std::size_t>(j) * rows_ + i]; }
  std::vector<double> col(int j) const { return std::vector<double>(d_.begin() + static_cast<std::size_t>(j) * rows_, d_.begin() + static_cast<std::size_t>(j + 1) * rows_); }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  void setZero() { std::fill(d_.begin(), d_.end(), 0.0); }
  bool operator==(const Matrix& o) const { return rows_ == o.rows_ && cols_ == o.cols_ && d_ == o.d_; }

 private:
  int rows_ = 0, cols_ = 0;
  std::vector<double> d_;
};
using StateTrajectory = Matrix;    // n x (T+1)
using ControlTrajectory = Matrix;  // m x T

// C ABI error code -> the exception type the reference would throw (SURVEY 8b error conventions).
inline void check(int rc) {
  if (rc == MAS_B200_OK) return;
  const std::string msg = mas_b200_last_error();
  if (rc == MAS_B200_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  if (rc == MAS_B200_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
  throw std::runtime_error(msg);
}

// Process-wide default device context (device 0 unless set_default_device is called first).
class Device {
 public:
  static int& default_device_id() {
    static int id = 0;
    return id;
  }
  static mas_b200_context_t context() {
    static Device dev;
    return dev.ctx_;
  }

 private:
  Device() { check(mas_b200_context_create(default_device_id(), nullptr, &ctx_)); }
  ~Device() { mas_b200_context_destroy(ctx_); }
  mas_b200_context_t ctx_ = nullptr;
};

// ---- ocp.hpp:30-237 -----------------------------------------------------------------------------------
struct OCP {
  StateTrajectory initial_states;
  ControlTrajectory initial_controls;
  StateTrajectory best_states;
  ControlTrajectory best_controls;
  double best_cost = std::numeric_limits<double>::max();

  State initial_state;
  int control_dim = 0;
  int state_dim = 0;
  int horizon_steps = 0;
  double dt = 0.0;

  std::optional<State> state_lower_bounds, state_upper_bounds;  // kept for interface parity; iLQR ignores them
  std::optional<Control> input_lower_bounds, input_upper_bounds;

  std::size_t id = 0;

  // In place of the std::function callbacks (types.hpp:21-50): a registered device model, its
  // constants, and which derivative callbacks are analytic (the rest are the FD defaults of ocp.hpp:117-135).
  int model_id = -1;
  std::vector<double> model_params;
  unsigned deriv_mask = 0;

  // counters of the last solve (the reference's solve() returns nothing)
  int last_iterations = 0;
  int last_status = MAS_B200_STATUS_MAX_ITER;

  mas_b200_ocp_desc desc() const {
    mas_b200_ocp_desc d;
    std::memset(&d, 0, sizeof(d));
    d.model_id = model_id;
    d.state_dim = state_dim;
    d.control_dim = control_dim;
    d.horizon_steps = horizon_steps;
    d.dt = dt;
    d.deriv_mask = deriv_mask;
    d.has_input_bounds = (input_lower_bounds && input_upper_bounds) ? 1 : 0;  // ilqr.hpp:213
    if (control_dim > MAS_B200_MAX_CONTROL_DIM) throw std::invalid_argument("control_dim exceeds MAS_B200_MAX_CONTROL_DIM");
    if (d.has_input_bounds)
      for (int i = 0; i < control_dim; ++i) {
        d.input_lower[i] = (*input_lower_bounds)[i];
        d.input_upper[i] = (*input_upper_bounds)[i];
      }
    if (model_params.size() > MAS_B200_MAX_PARAMS) throw std::invalid_argument("too many model parameters");
    d.num_params = static_cast<int>(model_params.size());
    for (std::size_t i = 0; i < model_params.size(); ++i) d.params[i] = model_params[i];
    return d;
  }

  // ocp.hpp:102-183: default controls, initial rollout, best_* and best_cost
  void initialize_problem() {
    if (initial_controls.rows() != control_dim || initial_controls.cols() != horizon_steps)
      initial_controls = ControlTrajectory::Zero(control_dim, horizon_steps);
    rollout(initial_controls, initial_states, best_cost);
    best_states = initial_states;
    best_controls = initial_controls;
  }

  // ocp.hpp:83-93
  void reset() {
    initial_controls = ControlTrajectory::Zero(control_dim, horizon_steps);
    rollout(initial_controls, initial_states, best_cost);
    best_states = initial_states;
    best_controls = initial_controls;
  }

  // ocp.hpp:95-100
  void update_initial_with_best() {
    initial_controls = best_controls;
    initial_states = best_states;
  }

  // ocp.hpp:186-236 (the reference's asserts vanish under NDEBUG; these throw instead)
  bool verify_problem() const {
    if (state_dim == 0 || control_dim == 0 || horizon_steps == 0 || dt == 0.0) throw std::invalid_argument("OCP dimensions / dt not set");
    if (static_cast<int>(initial_state.size()) != state_dim) throw std::invalid_argument("Initial state size does not match state dimension");
    if (input_lower_bounds && static_cast<int>(input_lower_bounds->size()) != control_dim) throw std::invalid_argument("Input lower bounds size mismatch");
    if (input_upper_bounds && static_cast<int>(input_upper_bounds->size()) != control_dim) throw std::invalid_argument("Input upper bounds size mismatch");
    return true;
  }

  // integrate_horizon + objective (ocp.hpp:110-113) on the device
  void rollout(const ControlTrajectory& U, StateTrajectory& X, double& cost) const {
    verify_problem();
    const mas_b200_ocp_desc d = desc();
    mas_b200_batch_t b = nullptr;
    check(mas_b200_batch_create(Device::context(), &d, 1, &b));
    struct Guard {
      mas_b200_batch_t b;
      ~Guard() { mas_b200_batch_destroy(b); }
    } guard{b};
    check(mas_b200_batch_set_initial_states(b, initial_state.data()));
    check(mas_b200_batch_set_controls(b, U.data()));
    check(mas_b200_batch_initialize(b));
    X = StateTrajectory(state_dim, horizon_steps + 1);
    check(mas_b200_batch_get_solution(b, X.data(), nullptr, &cost, nullptr, nullptr));
  }
};

// ---- solvers/ilqr.hpp:23-55 ------------------------------------------------------------------------
class iLQR {
 public:
  iLQR() { mas_b200_ilqr_default_params(&p_); }

  // .at() on the three required keys throws std::out_of_range exactly like the reference (ilqr.hpp:42-44)
  void set_params(const SolverParams& params) {
    p_.max_iterations = static_cast<int>(params.at("max_iterations"));
    p_.tolerance = params.at("tolerance");
    p_.max_ms = params.at("max_ms");
    p_.debug = params.count("debug") && params.at("debug") > 0.5;
    if (auto it = params.find("penalty"); it != params.end()) p_.penalty = it->second;
    if (auto it = params.find("penalty_increase"); it != params.end()) p_.penalty_increase = it->second;
    if (auto it = params.find("constraint_tolerance"); it != params.end()) p_.constraint_tolerance = it->second;
    if (auto it = params.find("inequality_activation_tolerance"); it != params.end()) p_.inequality_activation_tolerance = it->second;
  }

  // solve one problem in place: warm-starts from best_controls, writes best_states / best_controls / best_cost
  void solve(OCP& problem) {
    std::vector<OCP*> one{&problem};
    solve_batch(one);
  }

  // Batched entry the reference lacks: all problems must share dims, horizon, dt, bounds, model and
  // derivative mode (model constants may differ per problem).
  void solve_batch(const std::vector<OCP*>& problems) {
    if (problems.empty()) return;
    const OCP& first = *problems.front();
    first.verify_problem();
    mas_b200_ocp_desc d = first.desc();
    const int n = first.state_dim, m = first.control_dim, T = first.horizon_steps, B = static_cast<int>(problems.size());
    const int np = d.num_params;
    std::vector<double> x0(static_cast<std::size_t>(B) * n), U(static_cast<std::size_t>(B) * m * T), X(static_cast<std::size_t>(B) * n * (T + 1)),
        cost(B), prm(static_cast<std::size_t>(B) * np);
    std::vector<int> it(B), st(B);
    for (int b = 0; b < B; ++b) {
      const OCP& o = *problems[b];
      if (o.state_dim != n || o.control_dim != m || o.horizon_steps != T || o.dt != first.dt || o.model_id != first.model_id ||
          o.deriv_mask != first.deriv_mask || static_cast<int>(o.model_params.size()) != np)
        throw std::invalid_argument("solve_batch: problems must share shape, model and derivative mode");
      std::copy(o.initial_state.begin(), o.initial_state.end(), x0.begin() + static_cast<std::size_t>(b) * n);
      if (o.best_controls.rows() != m || o.best_controls.cols() != T) throw std::invalid_argument("best_controls has the wrong shape; call initialize_problem()");
      std::copy(o.best_controls.data(), o.best_controls.data() + static_cast<std::size_t>(m) * T, U.begin() + static_cast<std::size_t>(b) * m * T);
      std::copy(o.model_params.begin(), o.model_params.end(), prm.begin() + static_cast<std::size_t>(b) * np);
    }
    check(mas_b200_ilqr_solve_batch(Device::context(), &d, &p_, B, x0.data(), np > 0 ? prm.data() : nullptr, U.data(), X.data(), cost.data(),
                                    it.data(), st.data()));
    for (int b = 0; b < B; ++b) {
      OCP& o = *problems[b];
      o.best_states = StateTrajectory(n, T + 1);
      std::copy(X.begin() + static_cast<std::size_t>(b) * n * (T + 1), X.begin() + static_cast<std::size_t>(b + 1) * n * (T + 1), o.best_states.data());
      std::copy(U.begin() + static_cast<std::size_t>(b) * m * T, U.begin() + static_cast<std::size_t>(b + 1) * m * T, o.best_controls.data());
      o.best_cost = cost[b];
      o.last_iterations = it[b];
      o.last_status = st[b];
    }
    if (p_.debug) {  // the lines the reference prints while solving (ilqr.hpp:79-80,262-267), from the recorded trace
      std::vector<double> rec(static_cast<std::size_t>(p_.max_iterations + 1) * 6);
      for (int b = 0; b < B; ++b) {
        int nrec = 0;
        check(mas_b200_ilqr_last_debug_trace(Device::context(), b, p_.max_iterations + 1, rec.data(), &nrec));
        if (nrec > 0) std::cout << "iLQR initial cost=" << rec[0] << " merit=" << rec[1] << '\n';
        for (int r = 1; r < nrec; ++r)
          std::cout << "iLQR iter " << (r - 1) << ": cost=" << rec[6 * r + 0] << " merit=" << rec[6 * r + 1] << " d_merit=" << rec[6 * r + 2]
                    << " eq_violation=" << rec[6 * r + 3] << " ineq_violation=" << rec[6 * r + 4] << '\n';
      }
    }
  }

  const mas_b200_ilqr_params& raw_params() const { return p_; }

 private:
  mas_b200_ilqr_params p_;
};

// ---- solvers/solver.hpp:17-45 -------------------------------------------------------------------------
using Solver = std::variant<iLQR>;  // the reference's variant also holds CGD / OSQP: different algorithms, not on this path

inline void solve(Solver& solver, OCP& problem) {
  std::visit([&](auto& s) { s.solve(problem); }, solver);
}
inline void set_params(Solver& solver, const SolverParams& params) {
  std::visit([&](auto& s) { s.set_params(params); }, solver);
}
template <typename SolverT>
std::shared_ptr<Solver> create() {
  return std::make_shared<Solver>(std::in_place_type<SolverT>);
}

// ---- agent.hpp:9-44, solution.hpp:9-15 -------------------------------------------------------------------
struct Agent {
  std::size_t id;
  std::shared_ptr<OCP> ocp;
  Agent(std::size_t id_, std::shared_ptr<OCP> ocp_) : id(id_), ocp(std::move(ocp_)) {}
  int state_dim() const { return ocp->state_dim; }
  int control_dim() const { return ocp->control_dim; }
  void reset() { ocp->reset(); }
  void update_initial_with_best() { ocp->update_initial_with_best(); }
};
using AgentPtr = std::shared_ptr<Agent>;

struct Solution {
  std::vector<StateTrajectory> states;
  std::vector<ControlTrajectory> controls;
  std::vector<double> costs;
  double total_cost = 0.0;
};

// ---- multi_agent_problem.hpp:15-50 ---------------------------------------------------------------------
struct AgentBlockInfo {
  std::size_t agent_id;
  int state_offset, control_offset, state_dim, control_dim;
  AgentPtr agent;
};

class MultiAgentProblem {
 public:
  std::vector<AgentPtr> agents;
  std::vector<AgentBlockInfo> blocks;
  void add_agent(const AgentPtr& a) { agents.push_back(a); }
  void compute_offsets() {  // blocks sorted by agent id (:37-50)
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
};

namespace detail {

// Runs one strategy on the device for the agents of `problem` (one scenario) and writes every agent's
// best_* back, then update_initial_with_best (nash.hpp:66-71,232,241).
inline Solution run_strategy(int kind, int max_outer, const mas_b200_ilqr_params& prm, MultiAgentProblem& problem) {
  problem.compute_offsets();
  Solution sol;
  if (problem.blocks.empty()) return sol;
  const OCP& first = *problem.blocks.front().agent->ocp;
  mas_b200_ocp_desc d = first.desc();
  const int n = first.state_dim, m = first.control_dim, T = first.horizon_steps, A = static_cast<int>(problem.blocks.size());
  const int np = d.num_params;
  bool mixed = false;
  for (int a = 0; a < A; ++a) {
    const OCP& o = *problem.blocks[a].agent->ocp;
    mixed = mixed || o.state_dim != n || o.control_dim != m || o.horizon_steps != T || o.dt != first.dt || o.model_id != first.model_id ||
            o.deriv_mask != first.deriv_mask || static_cast<int>(o.model_params.size()) != np;
  }
  if (mixed) {
    // agents of different models / shapes (MultiAgentProblem accepts any mix): one description per agent, per-agent arrays
    // (mas_b200_strategy_run_mixed); the centralized strategy returns every agent's rows of the stacked solution, whose
    // horizon is the first block's (multi_agent_problem.hpp:65-69, centralized.hpp:27-36)
    const bool central = kind == MAS_B200_STRATEGY_CENTRALIZED;
    std::vector<mas_b200_ocp_desc> descs(A);
    std::vector<std::vector<double>> vx0(A), vp(A), vU0(A), vX(A), vU(A);
    std::vector<const double*> px0(A), pp(A), pU0(A);
    std::vector<double*> pX(A), pU(A), pc(A);
    std::vector<double> c(A);
    for (int a = 0; a < A; ++a) {
      const OCP& o = *problem.blocks[a].agent->ocp;
      descs[a] = o.desc();
      vx0[a].assign(o.initial_state.begin(), o.initial_state.end());
      vp[a].assign(o.model_params.begin(), o.model_params.end());
      if (o.best_controls.rows() != o.control_dim || o.best_controls.cols() != o.horizon_steps)
        throw std::invalid_argument("best_controls has the wrong shape; call initialize_problem()");
      vU0[a].assign(o.best_controls.data(), o.best_controls.data() + static_cast<std::size_t>(o.control_dim) * o.horizon_steps);
      const int Tr = central ? T : o.horizon_steps;
      vX[a].resize(static_cast<std::size_t>(o.state_dim) * (Tr + 1));
      vU[a].resize(static_cast<std::size_t>(o.control_dim) * Tr);
      px0[a] = vx0[a].data();
      pp[a] = vp[a].empty() ? nullptr : vp[a].data();
      pU0[a] = vU0[a].data();
      pX[a] = vX[a].data();
      pU[a] = vU[a].data();
      pc[a] = &c[a];
    }
    double total = 0.0;
    check(mas_b200_strategy_run_mixed(Device::context(), kind, descs.data(), &prm, max_outer, 1, A, px0.data(), pp.data(), pU0.data(), pX.data(),
                                      pU.data(), pc.data(), &total, nullptr));
    for (int a = 0; a < A; ++a) {
      OCP& o = *problem.blocks[a].agent->ocp;
      const int Tr = central ? T : o.horizon_steps;
      o.best_states = StateTrajectory(o.state_dim, Tr + 1);
      o.best_controls = ControlTrajectory(o.control_dim, Tr);
      std::copy(vX[a].begin(), vX[a].end(), o.best_states.data());
      std::copy(vU[a].begin(), vU[a].end(), o.best_controls.data());
      o.best_cost = c[a];
      if (!central) o.update_initial_with_best();  // the centralized strategy does not (centralized.hpp:18-38)
      sol.states.push_back(o.best_states);
      sol.controls.push_back(o.best_controls);
      sol.costs.push_back(o.best_cost);
    }
    sol.total_cost = total;
    return sol;
  }
  std::vector<double> x0(static_cast<std::size_t>(A) * n), prms(static_cast<std::size_t>(A) * np), X(static_cast<std::size_t>(A) * n * (T + 1)),
      U(static_cast<std::size_t>(A) * m * T), U0(static_cast<std::size_t>(A) * m * T), costs(A);
  for (int a = 0; a < A; ++a) {
    const OCP& o = *problem.blocks[a].agent->ocp;
    std::copy(o.initial_state.begin(), o.initial_state.end(), x0.begin() + static_cast<std::size_t>(a) * n);
    std::copy(o.model_params.begin(), o.model_params.end(), prms.begin() + static_cast<std::size_t>(a) * np);
    if (o.best_controls.rows() != m || o.best_controls.cols() != T) throw std::invalid_argument("best_controls has the wrong shape; call initialize_problem()");
    std::copy(o.best_controls.data(), o.best_controls.data() + static_cast<std::size_t>(m) * T, U0.begin() + static_cast<std::size_t>(a) * m * T);
  }
  double total = 0.0;
  // centralized: the stacked problem starts from zero controls in the reference (build_global_ocp never sets
  // initial_controls), so the agents' best_controls are not passed
  const double* u_init = kind == MAS_B200_STRATEGY_CENTRALIZED ? nullptr : U0.data();
  check(mas_b200_strategy_run(Device::context(), kind, &d, &prm, max_outer, 1, A, x0.data(), np > 0 ? prms.data() : nullptr, u_init, X.data(),
                              U.data(), costs.data(), &total, nullptr, nullptr, nullptr));
  for (int a = 0; a < A; ++a) {
    OCP& o = *problem.blocks[a].agent->ocp;
    o.best_states = StateTrajectory(n, T + 1);
    o.best_controls = ControlTrajectory(m, T);
    std::copy(X.begin() + static_cast<std::size_t>(a) * n * (T + 1), X.begin() + static_cast<std::size_t>(a + 1) * n * (T + 1), o.best_states.data());
    std::copy(U.begin() + static_cast<std::size_t>(a) * m * T, U.begin() + static_cast<std::size_t>(a + 1) * m * T, o.best_controls.data());
    o.best_cost = costs[a];
    // the Nash strategies end every accepted / restored round with update_initial_with_best (nash.hpp:66-71,232,241);
    // CentralizedStrategy writes best_* only (centralized.hpp:29-36)
    if (kind != MAS_B200_STRATEGY_CENTRALIZED) o.update_initial_with_best();
    sol.states.push_back(o.best_states);
    sol.controls.push_back(o.best_controls);
    sol.costs.push_back(o.best_cost);
  }
  sol.total_cost = total;
  return sol;
}

inline mas_b200_ilqr_params params_of(const Solver& proto, const SolverParams* params) {
  // nash.hpp:17-21,82-83: clones are default-constructed and then given `params`
  iLQR fresh;
  if (params) fresh.set_params(*params);
  else fresh = std::get<iLQR>(proto);
  return fresh.raw_params();
}

}  // namespace detail

// ---- strategies/centralized.hpp:10-39, strategies/nash.hpp:252-307 ------------------------------------------
struct CentralizedStrategy {
  Solver solver;
  explicit CentralizedStrategy(Solver s) : solver(std::move(s)) {}
  Solution operator()(MultiAgentProblem& problem) {
    return detail::run_strategy(MAS_B200_STRATEGY_CENTRALIZED, 1, detail::params_of(solver, nullptr), problem);
  }
};
struct SequentialNashStrategy {
  int max_outer;
  Solver solver_proto;
  SolverParams params;
  SequentialNashStrategy(int outer, Solver s, SolverParams p) : max_outer(outer), solver_proto(std::move(s)), params(std::move(p)) {}
  Solution operator()(MultiAgentProblem& problem) {
    return detail::run_strategy(MAS_B200_STRATEGY_SEQUENTIAL, max_outer, detail::params_of(solver_proto, &params), problem);
  }
};
struct LineSearchNashStrategy {
  int max_outer;
  Solver solver_proto;
  SolverParams params;
  LineSearchNashStrategy(int outer, Solver s, SolverParams p) : max_outer(outer), solver_proto(std::move(s)), params(std::move(p)) {}
  Solution operator()(MultiAgentProblem& problem) {
    return detail::run_strategy(MAS_B200_STRATEGY_LINESEARCH, max_outer, detail::params_of(solver_proto, &params), problem);
  }
};
struct TrustRegionNashStrategy {
  int max_outer;
  Solver solver_proto;
  SolverParams params;
  TrustRegionNashStrategy(int outer, Solver s, SolverParams p) : max_outer(outer), solver_proto(std::move(s)), params(std::move(p)) {}
  Solution operator()(MultiAgentProblem& problem) {
    return detail::run_strategy(MAS_B200_STRATEGY_TRUSTREGION, max_outer, detail::params_of(solver_proto, &params), problem);
  }
};

// ---- strategies/strategy.hpp:13-19 ------------------------------------------------------------------------
using Strategy = std::variant<CentralizedStrategy, SequentialNashStrategy, LineSearchNashStrategy, TrustRegionNashStrategy>;
inline Solution solve(Strategy& strategy, MultiAgentProblem& problem) {
  return std::visit([&](auto& s) { return s(problem); }, strategy);
}

// ---- examples/example_utils.hpp:19-110: name registries ------------------------------------------------------
namespace registry {

inline std::string normalize_key(const std::string& value) {
  std::string out;
  for (unsigned char ch : value)
    if (std::isalnum(ch)) out.push_back(static_cast<char>(std::tolower(ch)));
  return out;
}
inline std::string canonical_solver_name(const std::string& name) {
  const std::string k = normalize_key(name);
  if (k == "ilqr" || k == "primaldualilqr" || k == "pdilqr") return "ilqr";
  throw std::invalid_argument("Unknown solver '" + name + "'.");  // cgd / osqp live off the device path
}
inline std::string canonical_strategy_name(const std::string& name) {
  const std::string k = normalize_key(name);
  if (k == "centralized" || k == "centralised") return "centralized";
  if (k == "sequential" || k == "sequentialnash") return "sequential";
  if (k == "linesearch" || k == "linesearchnash") return "linesearch";
  if (k == "trustregion" || k == "trustregionnash") return "trustregion";
  throw std::invalid_argument("Unknown strategy '" + name + "'.");
}
inline Solver make_solver(const std::string& name) {
  canonical_solver_name(name);
  return Solver{std::in_place_type<iLQR>};
}
inline Strategy make_strategy(const std::string& name, Solver solver, const SolverParams& params, int max_outer) {
  const std::string c = canonical_strategy_name(name);
  if (c == "centralized") {
    set_params(solver, params);
    return Strategy{CentralizedStrategy{std::move(solver)}};
  }
  if (c == "sequential") return Strategy{SequentialNashStrategy{max_outer, std::move(solver), params}};
  if (c == "linesearch") return Strategy{LineSearchNashStrategy{max_outer, std::move(solver), params}};
  return Strategy{TrustRegionNashStrategy{max_outer, std::move(solver), params}};
}

}  // namespace registry

// ---- the example OCPs as the reference mains build them --------------------------------------------------------
namespace examples {

inline OCP from_example(int model_id, const State& x0, int horizon_steps = 0) {
  mas_b200_ocp_desc d;
  check(mas_b200_example_desc(model_id, &d));
  OCP p;
  p.model_id = model_id;
  p.state_dim = d.state_dim;
  p.control_dim = d.control_dim;
  p.horizon_steps = horizon_steps > 0 ? horizon_steps : d.horizon_steps;
  p.dt = d.dt;
  p.deriv_mask = d.deriv_mask;
  p.initial_state = x0;
  p.model_params.assign(d.params, d.params + d.num_params);
  if (model_id == MAS_B200_MODEL_PENDULUM) p.model_params[0] = static_cast<double>(p.horizon_steps);
  if (d.has_input_bounds) {
    p.input_lower_bounds = Control(d.input_lower, d.input_lower + d.control_dim);
    p.input_upper_bounds = Control(d.input_upper, d.input_upper + d.control_dim);
  }
  p.initial_controls = ControlTrajectory(p.control_dim, p.horizon_steps);
  check(mas_b200_example_controls(model_id, p.horizon_steps, p.initial_controls.data()));
  p.initialize_problem();
  p.verify_problem();
  return p;
}
// examples/single_track_ocp.cpp:14-116
inline OCP create_single_track_lane_following_ocp() { return from_example(MAS_B200_MODEL_SINGLE_TRACK_LANE, {0.0, 1.0, 0.0, 0.0}); }
// examples/multi_agent_single_track.cpp:31-72
inline OCP create_single_track_circular_ocp(double initial_theta, double track_radius, double target_velocity, int time_steps) {
  mas_b200_ocp_desc d;
  check(mas_b200_example_desc(MAS_B200_MODEL_SINGLE_TRACK_CIRC, &d));
  OCP p = from_example(MAS_B200_MODEL_SINGLE_TRACK_CIRC,
                       {track_radius * std::cos(initial_theta), track_radius * std::sin(initial_theta), 1.57 + initial_theta, 4.0}, time_steps);
  p.model_params[0] = track_radius;
  p.model_params[1] = target_velocity;
  p.initialize_problem();
  return p;
}
// examples/multi_agent_lqr.cpp:21-76 (n_x = n_u = 4 is the registered instance)
inline OCP create_linear_lqr_ocp(int n_x, int n_u, double dt, int T) {
  if (n_x != 4 || n_u != 4) throw std::invalid_argument("only the 4x4 LQR example is registered as a device model");
  OCP p = from_example(MAS_B200_MODEL_LQR4, {1.0, 0.0, 0.0, 0.0}, T);
  p.dt = dt;
  p.initialize_problem();
  return p;
}
// examples/pendulum_swing_up.cpp:29-117
inline OCP create_pendulum_swingup_ocp() { return from_example(MAS_B200_MODEL_PENDULUM, {M_PI - 0.05, 0.0}); }
// examples/rocket_max_altitude.cpp:31-137
inline OCP create_max_altitude_rocket_ocp() { return from_example(MAS_B200_MODEL_ROCKET, {0.0, 0.0, 1.0}); }

}  // namespace examples
}  // namespace mas_b200

#ifdef MAS_B200_NAMESPACE_ALIAS_MAS
namespace mas = mas_b200;
#endif
