// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).  Pinned bit for bit to a build of the reference's own sources
// (oracle/_ref, tests/test_ref_pin.py; see dense.hpp).
//
// ref_core.hpp: CPU restatement of the reference's problem model and numerical kernels, in the
// reference's own shape (std::function callbacks over dynamically sized vectors):
//   types.hpp:14-57            -> callback typedefs, SolverParams
//   integrator.hpp:19-48       -> integrate_rk4, integrate_horizon
//   finite_differences.hpp:53-287 -> central-difference Jacobians / gradients / Hessians / cross term
//   constraint_helpers.hpp:107-114 -> clamp_controls
//   ocp.hpp:14-28,30-237       -> compute_trajectory_cost, struct OCP, initialize_problem
#pragma once
#include <cmath>
#include <functional>
#include <limits>
#include <optional>
#include <string>
#include <unordered_map>

#include "dense.hpp"

namespace oracle {

using State = Vec;
using Control = Vec;
using StateTrajectory = Mat;    // n x (T+1), column t = x_t
using ControlTrajectory = Mat;  // m x T

// types.hpp:21-57
using MotionModel = std::function<Vec(const State&, const Control&)>;
using ObjectiveFunction = std::function<double(const StateTrajectory&, const ControlTrajectory&)>;
using StageCostFunction = std::function<double(const State&, const Control&, std::size_t)>;
using TerminalCostFunction = std::function<double(const State&)>;
using ConstraintsFunction = std::function<Vec(const State&, const Control&)>;
using ConstraintsJacobianFunction = std::function<Mat(const State&, const Control&)>;
using DynamicsJacobianFn = std::function<Mat(const MotionModel&, const State&, const Control&)>;
using CostGradientFn = std::function<Vec(const StageCostFunction&, const State&, const Control&, std::size_t)>;
using CostHessianFn = std::function<Mat(const StageCostFunction&, const State&, const Control&, std::size_t)>;
using TerminalGradientFn = std::function<Vec(const TerminalCostFunction&, const State&)>;
using TerminalHessianFn = std::function<Mat(const TerminalCostFunction&, const State&)>;
using SolverParams = std::unordered_map<std::string, double>;

// ---- integrator.hpp:19-28 ------------------------------------------------------------------
// x+ = x + (dt/6.0) * (((k1 + 2*k2) + 2*k3) + k4), stage points x + (0.5*dt)*k1 etc.
inline State integrate_rk4(const State& x, const Control& u, double dt, const MotionModel& f) {
  const Vec k1 = f(x, u);
  const Vec k2 = f(add(x, scale(0.5 * dt, k1)), u);
  const Vec k3 = f(add(x, scale(0.5 * dt, k2)), u);
  const Vec k4 = f(add(x, scale(dt, k3)), u);
  const Vec sum = add(add(add(k1, scale(2.0, k2)), scale(2.0, k3)), k4);
  return add(x, scale(dt / 6.0, sum));
}

// ---- integrator.hpp:31-48 ------------------------------------------------------------------
inline StateTrajectory integrate_horizon(const State& x0, const ControlTrajectory& U, double dt, const MotionModel& f) {
  StateTrajectory X(static_cast<int>(x0.size()), U.cols + 1);
  X.set_col(0, x0);
  State s = x0;
  for (int i = 0; i < U.cols; ++i) {
    s = integrate_rk4(s, U.col(i), dt, f);
    X.set_col(i + 1, s);
  }
  return X;
}

// ---- finite_differences.hpp:53-92 ------------------------------------------------------------
inline Mat compute_dynamics_state_jacobian(const MotionModel& f, const State& x, const Control& u) {
  const int n = static_cast<int>(x.size());
  const double eps = 1e-6;
  Mat A(n, n);
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    const Vec fp = f(add(x, dx), u);
    const Vec fm = f(sub(x, dx), u);
    for (int r = 0; r < n; ++r) A(r, i) = (fp[r] - fm[r]) / (2 * eps);
  }
  return A;
}

inline Mat compute_dynamics_control_jacobian(const MotionModel& f, const State& x, const Control& u) {
  const int n = static_cast<int>(x.size());
  const int m = static_cast<int>(u.size());
  const double eps = 1e-6;
  Mat B(n, m);
  for (int i = 0; i < m; ++i) {
    Vec du = zeros(m);
    du[i] = eps;
    const Vec fp = f(x, add(u, du));
    const Vec fm = f(x, sub(u, du));
    for (int r = 0; r < n; ++r) B(r, i) = (fp[r] - fm[r]) / (2 * eps);
  }
  return B;
}

// finite_differences.hpp:95-107
inline double safe_eval(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const double v = c(x, u, t);
  return std::isfinite(v) ? v : 0.0;
}
inline double safe_eval_terminal(const TerminalCostFunction& c, const State& x) {
  const double v = c(x);
  return std::isfinite(v) ? v : 0.0;
}

// finite_differences.hpp:110-136  (no safe_eval on gradients)
inline Vec compute_cost_state_gradient(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const int n = static_cast<int>(x.size());
  Vec g = zeros(n);
  const double eps = 1e-6;
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    g[i] = (c(add(x, dx), u, t) - c(sub(x, dx), u, t)) / (2 * eps);
  }
  return g;
}
inline Vec compute_cost_control_gradient(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const int m = static_cast<int>(u.size());
  Vec g = zeros(m);
  const double eps = 1e-6;
  for (int i = 0; i < m; ++i) {
    Vec du = zeros(m);
    du[i] = eps;
    g[i] = (c(x, add(u, du), t) - c(x, sub(u, du), t)) / (2 * eps);
  }
  return g;
}

// finite_differences.hpp:138-174: diagonal (f+ - 2 f + f-)/(eps*eps), f re-evaluated per i; every
// ordered off-diagonal pair separately, (f++ - f+- - f-+ + f--)/(4*eps*eps).
inline Mat compute_cost_state_hessian(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const int n = static_cast<int>(x.size());
  Mat H(n, n);
  const double eps = 1e-5;
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    const double fp = safe_eval(c, add(x, dx), u, t);
    const double f0 = safe_eval(c, x, u, t);
    const double fm = safe_eval(c, sub(x, dx), u, t);
    H(i, i) = (fp - 2 * f0 + fm) / (eps * eps);
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j) {
        Vec di = zeros(n), dj = zeros(n);
        di[i] = eps;
        dj[j] = eps;
        const double fpp = safe_eval(c, add(add(x, di), dj), u, t);
        const double fpm = safe_eval(c, sub(add(x, di), dj), u, t);
        const double fmp = safe_eval(c, add(sub(x, di), dj), u, t);
        const double fmm = safe_eval(c, sub(sub(x, di), dj), u, t);
        H(i, j) = (fpp - fpm - fmp + fmm) / (4 * eps * eps);
      }
  return H;
}

// finite_differences.hpp:176-210
inline Mat compute_cost_control_hessian(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const int m = static_cast<int>(u.size());
  Mat H(m, m);
  const double eps = 1e-5;
  for (int i = 0; i < m; ++i) {
    Vec du = zeros(m);
    du[i] = eps;
    const double fp = safe_eval(c, x, add(u, du), t);
    const double f0 = safe_eval(c, x, u, t);
    const double fm = safe_eval(c, x, sub(u, du), t);
    H(i, i) = (fp - 2 * f0 + fm) / (eps * eps);
  }
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      if (i != j) {
        Vec di = zeros(m), dj = zeros(m);
        di[i] = eps;
        dj[j] = eps;
        const double fpp = safe_eval(c, x, add(add(u, di), dj), t);
        const double fpm = safe_eval(c, x, sub(add(u, di), dj), t);
        const double fmp = safe_eval(c, x, add(sub(u, di), dj), t);
        const double fmm = safe_eval(c, x, sub(sub(u, di), dj), t);
        H(i, j) = (fpp - fpm - fmp + fmm) / (4 * eps * eps);
      }
  return H;
}

// finite_differences.hpp:212-261
inline Vec compute_terminal_cost_gradient(const TerminalCostFunction& c, const State& x) {
  const int n = static_cast<int>(x.size());
  Vec g = zeros(n);
  const double eps = 1e-6;
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    g[i] = (c(add(x, dx)) - c(sub(x, dx))) / (2 * eps);
  }
  return g;
}
inline Mat compute_terminal_cost_hessian(const TerminalCostFunction& c, const State& x) {
  const int n = static_cast<int>(x.size());
  Mat H(n, n);
  const double eps = 1e-5;
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    const double fp = safe_eval_terminal(c, add(x, dx));
    const double f0 = safe_eval_terminal(c, x);
    const double fm = safe_eval_terminal(c, sub(x, dx));
    H(i, i) = (fp - 2 * f0 + fm) / (eps * eps);
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      if (i != j) {
        Vec di = zeros(n), dj = zeros(n);
        di[i] = eps;
        dj[j] = eps;
        const double fpp = safe_eval_terminal(c, add(add(x, di), dj));
        const double fpm = safe_eval_terminal(c, sub(add(x, di), dj));
        const double fmp = safe_eval_terminal(c, add(sub(x, di), dj));
        const double fmm = safe_eval_terminal(c, sub(sub(x, di), dj));
        H(i, j) = (fpp - fpm - fmp + fmm) / (4 * eps * eps);
      }
  return H;
}

// finite_differences.hpp:263-287: m x n, eps 1e-6; f_pm perturbs x by -eps and u by +eps.
inline Mat compute_cost_cross_term(const StageCostFunction& c, const State& x, const Control& u, std::size_t t) {
  const int m = static_cast<int>(u.size());
  const int n = static_cast<int>(x.size());
  Mat H(m, n);
  const double eps = 1e-6;
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      Vec du = zeros(m), dx = zeros(n);
      du[i] = eps;
      dx[j] = eps;
      const double fpp = safe_eval(c, add(x, dx), add(u, du), t);
      const double fpm = safe_eval(c, sub(x, dx), add(u, du), t);
      const double fmp = safe_eval(c, add(x, dx), sub(u, du), t);
      const double fmm = safe_eval(c, sub(x, dx), sub(u, du), t);
      H(i, j) = (fpp - fpm - fmp + fmm) / (4 * eps * eps);
    }
  return H;
}

// finite_differences.hpp:289-345 (constraint Jacobians; dormant in every example)
inline Mat compute_constraints_state_jacobian(const ConstraintsFunction& g, const State& x, const Control& u) {
  if (!g) return Mat();
  const Vec base = g(x, u);
  const int m = static_cast<int>(base.size());
  if (m == 0) return Mat();
  const int n = static_cast<int>(x.size());
  const double eps = 1e-6;
  Mat J(m, n);
  for (int i = 0; i < n; ++i) {
    Vec dx = zeros(n);
    dx[i] = eps;
    const Vec fp = g(add(x, dx), u);
    const Vec fm = g(sub(x, dx), u);
    for (int r = 0; r < m; ++r) J(r, i) = (fp[r] - fm[r]) / (2 * eps);
  }
  return J;
}
inline Mat compute_constraints_control_jacobian(const ConstraintsFunction& g, const State& x, const Control& u) {
  if (!g) return Mat();
  const Vec base = g(x, u);
  const int m = static_cast<int>(base.size());
  if (m == 0) return Mat();
  const int p = static_cast<int>(u.size());
  const double eps = 1e-6;
  Mat J(m, p);
  for (int i = 0; i < p; ++i) {
    Vec du = zeros(p);
    du[i] = eps;
    const Vec fp = g(x, add(u, du));
    const Vec fm = g(x, sub(u, du));
    for (int r = 0; r < m; ++r) J(r, i) = (fp[r] - fm[r]) / (2 * eps);
  }
  return J;
}

// ---- constraint_helpers.hpp:107-114: per column min(upper) then max(lower) ---------------------
inline void clamp_controls(ControlTrajectory& U, const Control& lo, const Control& hi) {
  for (int t = 0; t < U.cols; ++t)
    for (int i = 0; i < U.rows; ++i) {
      double v = U(i, t);
      v = (hi[i] < v) ? hi[i] : v;  // cwiseMin(upper)
      v = (lo[i] > v) ? lo[i] : v;  // cwiseMax(lower)
      U(i, t) = v;
    }
}

// ---- ocp.hpp:14-28 --------------------------------------------------------------------------
inline double compute_trajectory_cost(const StateTrajectory& X, const ControlTrajectory& U, const StageCostFunction& stage,
                                      const TerminalCostFunction& terminal) {
  const int T = U.cols;
  double cost = 0.0;
  for (int t = 0; t < T; ++t) cost += stage(X.col(t), U.col(t), static_cast<std::size_t>(t));
  cost += terminal(X.col(X.cols - 1));
  return cost;
}

// ---- ocp.hpp:30-237 --------------------------------------------------------------------------
struct OCP {
  StateTrajectory initial_states;
  ControlTrajectory initial_controls;
  StateTrajectory best_states;
  ControlTrajectory best_controls;
  double best_cost = std::numeric_limits<double>::max();

  State initial_state;
  MotionModel dynamics;
  StageCostFunction stage_cost = [](const State&, const Control&, std::size_t) { return 0.0; };
  TerminalCostFunction terminal_cost = [](const State&) { return 0.0; };
  ObjectiveFunction objective_function;

  int control_dim = 0;
  int state_dim = 0;
  int horizon_steps = 0;
  double dt = 0.0;

  std::optional<State> state_lower_bounds, state_upper_bounds;
  std::optional<Control> input_lower_bounds, input_upper_bounds;

  ConstraintsFunction equality_constraints;
  ConstraintsFunction inequality_constraints;
  ConstraintsJacobianFunction equality_constraints_state_jacobian, equality_constraints_control_jacobian;
  ConstraintsJacobianFunction inequality_constraints_state_jacobian, inequality_constraints_control_jacobian;

  DynamicsJacobianFn dynamics_state_jacobian, dynamics_control_jacobian;
  CostGradientFn cost_state_gradient, cost_control_gradient;
  CostHessianFn cost_state_hessian, cost_control_hessian, cost_cross_term;
  TerminalGradientFn terminal_cost_gradient;
  TerminalHessianFn terminal_cost_hessian;

  std::size_t id = 0;

  // ocp.hpp:83-93
  void reset() {
    initial_controls = ControlTrajectory(control_dim, horizon_steps);
    initial_states = integrate_horizon(initial_state, initial_controls, dt, dynamics);
    best_states = initial_states;
    best_controls = initial_controls;
    best_cost = objective_function(initial_states, initial_controls);
  }

  // ocp.hpp:95-100
  void update_initial_with_best() {
    initial_controls = best_controls;
    initial_states = best_states;
  }

  // ocp.hpp:102-183
  void initialize_problem() {
    if (initial_controls.rows != control_dim || initial_controls.cols != horizon_steps)
      initial_controls = ControlTrajectory(control_dim, horizon_steps);
    initial_states = integrate_horizon(initial_state, initial_controls, dt, dynamics);
    best_states = initial_states;
    best_controls = initial_controls;

    if (!dynamics_state_jacobian) dynamics_state_jacobian = compute_dynamics_state_jacobian;
    if (!dynamics_control_jacobian) dynamics_control_jacobian = compute_dynamics_control_jacobian;
    if (!cost_state_gradient) cost_state_gradient = compute_cost_state_gradient;
    if (!cost_control_gradient) cost_control_gradient = compute_cost_control_gradient;
    if (!cost_state_hessian) cost_state_hessian = compute_cost_state_hessian;
    if (!cost_control_hessian) cost_control_hessian = compute_cost_control_hessian;
    if (!cost_cross_term) cost_cross_term = compute_cost_cross_term;
    if (!terminal_cost_gradient) terminal_cost_gradient = compute_terminal_cost_gradient;
    if (!terminal_cost_hessian) terminal_cost_hessian = compute_terminal_cost_hessian;

    if (equality_constraints) {
      if (!equality_constraints_state_jacobian) {
        auto g = equality_constraints;
        equality_constraints_state_jacobian = [g](const State& x, const Control& u) { return compute_constraints_state_jacobian(g, x, u); };
      }
      if (!equality_constraints_control_jacobian) {
        auto g = equality_constraints;
        equality_constraints_control_jacobian = [g](const State& x, const Control& u) { return compute_constraints_control_jacobian(g, x, u); };
      }
    }
    if (inequality_constraints) {
      if (!inequality_constraints_state_jacobian) {
        auto g = inequality_constraints;
        inequality_constraints_state_jacobian = [g](const State& x, const Control& u) { return compute_constraints_state_jacobian(g, x, u); };
      }
      if (!inequality_constraints_control_jacobian) {
        auto g = inequality_constraints;
        inequality_constraints_control_jacobian = [g](const State& x, const Control& u) { return compute_constraints_control_jacobian(g, x, u); };
      }
    }

    if (!objective_function) {
      auto sc = stage_cost;
      auto tc = terminal_cost;
      objective_function = [sc, tc](const StateTrajectory& X, const ControlTrajectory& U) { return compute_trajectory_cost(X, U, sc, tc); };
    }
    best_cost = objective_function(initial_states, initial_controls);
  }
};

}  // namespace oracle
