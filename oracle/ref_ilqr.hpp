// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).  Pinned bit for bit to a build of the reference's own sources
// (oracle/_ref, tests/test_ref_pin.py; see dense.hpp).
//
// ref_ilqr.hpp: CPU restatement of the reference's augmented-Lagrangian iLQR,
// include/multi_agent_solver/solvers/ilqr.hpp:26-55 (params), :59-273 (solve), :278-377 (buffers),
// :380-407 (merit).  The reference returns nothing from solve(); the counters below are the
// oracle's definition: iterations = loop bodies that reached the backward pass,
// status = why the loop ended.
#pragma once
#include <chrono>
#include <cstdio>

#include "ref_core.hpp"

namespace oracle {

enum SolveStatus { STATUS_CONVERGED = 0, STATUS_MAX_ITER = 1, STATUS_TIME_LIMIT = 2 };

struct SolveStats {
  int iterations = 0;
  int status = STATUS_MAX_ITER;
  int rollouts = 0;        // forward rollouts incl. the prologue one
  int alpha_trials = 0;    // line-search candidates evaluated
  int reg_retries = 0;     // Q_uu + reg*I retries (ilqr.hpp:175-182)
  std::vector<double> cost_trace;   // cost after each iteration
  std::vector<int> alpha_index;     // accepted candidate index per iteration, -1 = none
};

struct OracleOptions {
  bool aliased_symmetrize = true;  // ilqr.hpp:102,192 as compiled (quirk 3); false = exact symmetrisation
};

class iLQR {
 public:
  // ilqr.hpp:26-37
  iLQR()
      : max_iterations(50), tolerance(1e-6), max_ms(std::numeric_limits<double>::infinity()), debug(false), penalty_parameter(10.0),
        penalty_increase(5.0), constraint_tolerance(1e-4), inequality_activation_tolerance(1e-6), equality_dim(0), inequality_dim(0) {}

  // ilqr.hpp:39-55 (.at() throws std::out_of_range for the three required keys)
  void set_params(const SolverParams& p) {
    max_iterations = static_cast<int>(p.at("max_iterations"));
    tolerance = p.at("tolerance");
    max_ms = p.at("max_ms");
    debug = p.count("debug") && p.at("debug") > 0.5;
    if (auto it = p.find("penalty"); it != p.end()) penalty_parameter = it->second;
    if (auto it = p.find("penalty_increase"); it != p.end()) penalty_increase = it->second;
    if (auto it = p.find("constraint_tolerance"); it != p.end()) constraint_tolerance = it->second;
    if (auto it = p.find("inequality_activation_tolerance"); it != p.end()) inequality_activation_tolerance = it->second;
  }

  OracleOptions options;
  SolveStats stats;  // of the last solve()

  // ilqr.hpp:59-273
  void solve(OCP& problem) {
    using clock = std::chrono::high_resolution_clock;
    const auto start = clock::now();
    stats = SolveStats();

    resize_buffers(problem);

    const int T = problem.horizon_steps;
    const int nx = problem.state_dim;
    const int nu = problem.control_dim;
    const double dt = problem.dt;

    StateTrajectory& x = problem.best_states;
    ControlTrajectory& u = problem.best_controls;
    double& cost = problem.best_cost;

    x = integrate_horizon(problem.initial_state, u, dt, problem.dynamics);  // :75
    stats.rollouts++;
    cost = problem.objective_function(x, u);                                  // :76
    double current_merit = compute_merit(problem, x, u);                     // :78
    if (debug) std::printf("iLQR initial cost=%.17g merit=%.17g\n", cost, current_merit);

    const Mat identity_nu = Mat::identity(nu);
    stats.status = STATUS_MAX_ITER;

    for (int iter = 0; iter < max_iterations; ++iter) {
      // :84-90 integer-millisecond budget, checked only here
      const double elapsed_ms =
          static_cast<double>(std::chrono::duration_cast<std::chrono::milliseconds>(clock::now() - start).count());
      if (elapsed_ms > max_ms) {
        stats.status = STATUS_TIME_LIMIT;
        break;
      }
      stats.iterations++;

      // :92-102
      Vec v_x = problem.terminal_cost_gradient ? problem.terminal_cost_gradient(problem.terminal_cost, x.col(T)) : zeros(nx);
      Mat v_xx = problem.terminal_cost_hessian ? problem.terminal_cost_hessian(problem.terminal_cost, x.col(T)) : Mat(nx, nx);
      symmetrize(v_xx);

      for (int t = T - 1; t >= 0; --t) {
        const Vec xt = x.col(t), ut = u.col(t);
        const std::size_t ti = static_cast<std::size_t>(t);
        // :106-113
        a_step[t] = problem.dynamics_state_jacobian(problem.dynamics, xt, ut);
        b_step[t] = problem.dynamics_control_jacobian(problem.dynamics, xt, ut);
        const Vec l_x = problem.cost_state_gradient(problem.stage_cost, xt, ut, ti);
        const Vec l_u = problem.cost_control_gradient(problem.stage_cost, xt, ut, ti);
        const Mat l_xx = problem.cost_state_hessian(problem.stage_cost, xt, ut, ti);
        const Mat l_uu = problem.cost_control_hessian(problem.stage_cost, xt, ut, ti);
        const Mat l_ux = problem.cost_cross_term(problem.stage_cost, xt, ut, ti);
        const Mat& A = a_step[t];
        const Mat& B = b_step[t];

        // :115-119   (A^T V_xx) and (B^T V_xx) are evaluated first, then multiplied on the right
        Vec q_x = add(l_x, matTvec(A, v_x));
        Vec q_u = add(l_u, matTvec(B, v_x));
        const Mat AtV = matTmul(A, v_xx);
        const Mat BtV = matTmul(B, v_xx);
        Mat q_xx = add(l_xx, matmul(AtV, A));
        Mat q_ux = add(l_ux, matmul(BtV, A));
        Mat q_uu = add(l_uu, matmul(BtV, B));

        // :121-141 equality AL terms
        if (equality_dim > 0 && problem.equality_constraints) {
          const Vec c = problem.equality_constraints(xt, ut);
          const Mat Jx = problem.equality_constraints_state_jacobian ? problem.equality_constraints_state_jacobian(xt, ut)
                                                                     : compute_constraints_state_jacobian(problem.equality_constraints, xt, ut);
          const Mat Ju = problem.equality_constraints_control_jacobian ? problem.equality_constraints_control_jacobian(xt, ut)
                                                                       : compute_constraints_control_jacobian(problem.equality_constraints, xt, ut);
          const Vec dual = add(eq_multipliers[t], scale(penalty_parameter, c));
          q_x = add(q_x, matTvec(Jx, dual));
          q_u = add(q_u, matTvec(Ju, dual));
          q_xx = add(q_xx, matmul(scale(penalty_parameter, transpose(Jx)), Jx));
          q_ux = add(q_ux, matmul(scale(penalty_parameter, transpose(Ju)), Jx));
          q_uu = add(q_uu, matmul(scale(penalty_parameter, transpose(Ju)), Ju));
        }

        // :143-170 inequality AL terms
        if (inequality_dim > 0 && problem.inequality_constraints) {
          const Vec g = problem.inequality_constraints(xt, ut);
          const Mat Jx = problem.inequality_constraints_state_jacobian ? problem.inequality_constraints_state_jacobian(xt, ut)
                                                                       : compute_constraints_state_jacobian(problem.inequality_constraints, xt, ut);
          const Mat Ju = problem.inequality_constraints_control_jacobian ? problem.inequality_constraints_control_jacobian(xt, ut)
                                                                         : compute_constraints_control_jacobian(problem.inequality_constraints, xt, ut);
          const int p = static_cast<int>(g.size());
          Vec dual(p);
          bool any_active = false;
          Vec active(p);
          for (int i = 0; i < p; ++i) {
            const double slack = g[i] > 0.0 ? g[i] : 0.0;
            active[i] = (g[i] > -inequality_activation_tolerance) ? 1.0 : 0.0;
            any_active = any_active || active[i] != 0.0;
            dual[i] = ineq_multipliers[t][i] * active[i] + penalty_parameter * slack * active[i];
          }
          q_x = add(q_x, matTvec(Jx, dual));
          q_u = add(q_u, matTvec(Ju, dual));
          if (any_active) {
            Mat D(p, p);
            for (int i = 0; i < p; ++i) D(i, i) = active[i];
            q_xx = add(q_xx, matmul(matmul(scale(penalty_parameter, transpose(Jx)), D), Jx));
            q_ux = add(q_ux, matmul(matmul(scale(penalty_parameter, transpose(Ju)), D), Jx));
            q_uu = add(q_uu, matmul(matmul(scale(penalty_parameter, transpose(Ju)), D), Ju));
          }
        }

        // :172-183 cumulative diagonal shifts 1e-6, 1e-5, ... until LLT succeeds; explicit inverse
        Mat q_uu_reg = q_uu;
        LLT llt;
        double reg = 1e-6;
        while (true) {
          if (llt.compute(q_uu_reg)) break;
          for (int i = 0; i < nu; ++i) q_uu_reg(i, i) += reg * 1.0;
          reg *= 10.0;
          stats.reg_retries++;
        }
        const Mat q_uu_inv = llt.solve(identity_nu);

        // :185-186 gains use the regularised inverse
        k[t] = matvec(neg(q_uu_inv), q_u);
        k_matrix[t] = matmul(neg(q_uu_inv), q_ux);
        const Mat& K = k_matrix[t];

        // :188-192 value update uses the UNregularised Q_uu
        const Mat KtQuu = matTmul(K, q_uu);
        v_x = add(add(add(q_x, matTvec(K, q_u)), matTvec(q_ux, k[t])), matvec(KtQuu, k[t]));
        v_xx = add(add(add(q_xx, matTmul(K, q_ux)), matTmul(q_ux, K)), matmul(KtQuu, K));
        symmetrize(v_xx);
      }

      // :195-228 backtracking line search on the merit, first improvement wins
      StateTrajectory x_trial(nx, T + 1);
      ControlTrajectory u_trial(nu, T);
      x_trial.set_col(0, problem.initial_state);

      const double amin = 1e-3;
      double alpha = 1.0;
      double best_merit = current_merit;
      StateTrajectory best_x = x;
      ControlTrajectory best_u = u;
      int accepted = -1, cand = 0;

      while (alpha >= amin) {
        for (int t = 0; t < T; ++t) {
          const Vec dx = sub(x_trial.col(t), x.col(t));
          const Vec Kdx = matvec(k_matrix[t], dx);
          for (int i = 0; i < nu; ++i) u_trial(i, t) = (u(i, t) + alpha * k[t][i]) + Kdx[i];
          if (problem.input_lower_bounds && problem.input_upper_bounds)
            clamp_controls(u_trial, *problem.input_lower_bounds, *problem.input_upper_bounds);
          x_trial.set_col(t + 1, integrate_rk4(x_trial.col(t), u_trial.col(t), dt, problem.dynamics));
        }
        stats.rollouts++;
        stats.alpha_trials++;
        const double trial_merit = compute_merit(problem, x_trial, u_trial);
        if (trial_merit < best_merit) {
          best_merit = trial_merit;
          best_x = x_trial;
          best_u = u_trial;
          accepted = cand;
          break;
        }
        alpha *= 0.5;
        cand++;
      }

      // :230-234
      const double improvement = current_merit - best_merit;
      x = best_x;
      u = best_u;
      cost = problem.objective_function(x, u);
      current_merit = best_merit;
      stats.cost_trace.push_back(cost);
      stats.alpha_index.push_back(accepted);

      // :236-260 multiplier / penalty update
      double eq_violation_norm = 0.0, ineq_violation_norm = 0.0;
      for (int t = 0; t < T; ++t) {
        if (equality_dim > 0 && problem.equality_constraints) {
          const Vec r = problem.equality_constraints(x.col(t), u.col(t));
          for (std::size_t i = 0; i < r.size(); ++i) eq_multipliers[t][i] += penalty_parameter * r[i];
          eq_violation_norm += dot(r, r);
        }
        if (inequality_dim > 0 && problem.inequality_constraints) {
          const Vec r = problem.inequality_constraints(x.col(t), u.col(t));
          Vec pos(r.size());
          for (std::size_t i = 0; i < r.size(); ++i) {
            pos[i] = r[i] > 0.0 ? r[i] : 0.0;
            const double v = ineq_multipliers[t][i] + penalty_parameter * pos[i];
            ineq_multipliers[t][i] = v > 0.0 ? v : 0.0;
          }
          ineq_violation_norm += dot(pos, pos);
        }
      }
      eq_violation_norm = std::sqrt(eq_violation_norm);
      ineq_violation_norm = std::sqrt(ineq_violation_norm);
      if (eq_violation_norm > constraint_tolerance || ineq_violation_norm > constraint_tolerance) penalty_parameter *= penalty_increase;

      if (debug)
        std::printf("iLQR iter %d: cost=%.17g merit=%.17g d_merit=%.17g eq_violation=%g ineq_violation=%g\n", iter, cost, current_merit,
                    improvement, eq_violation_norm, ineq_violation_norm);

      // :269-271
      if (improvement < tolerance && eq_violation_norm < constraint_tolerance && ineq_violation_norm < constraint_tolerance) {
        stats.status = STATUS_CONVERGED;
        break;
      }
    }
  }

 private:
  void symmetrize(Mat& m) const {
    if (options.aliased_symmetrize) {
      symmetrize_aliased(m);
    } else {
      const Mat mt = transpose(m);
      m = scale(0.5, add(m, mt));
    }
  }

  // ilqr.hpp:278-377: gains and Jacobian buffers are zeroed on every call; multipliers and the
  // penalty parameter persist across calls on the same solver object unless the dimension changes.
  void resize_buffers(const OCP& problem) {
    const int T = problem.horizon_steps;
    const int nx = problem.state_dim;
    const int nu = problem.control_dim;
    k.assign(T, zeros(nu));
    k_matrix.assign(T, Mat(nu, nx));
    a_step.assign(T, Mat(nx, nx));
    b_step.assign(T, Mat(nx, nu));

    Control default_control = zeros(nu);
    if (problem.initial_controls.cols == T) default_control = problem.initial_controls.col(0);
    equality_dim = 0;
    inequality_dim = 0;
    if (problem.equality_constraints) equality_dim = static_cast<int>(problem.equality_constraints(problem.initial_state, default_control).size());
    if (problem.inequality_constraints)
      inequality_dim = static_cast<int>(problem.inequality_constraints(problem.initial_state, default_control).size());

    if (equality_dim > 0) {
      if (static_cast<int>(eq_multipliers.size()) != T) eq_multipliers.assign(T, zeros(equality_dim));
      else
        for (auto& m : eq_multipliers)
          if (static_cast<int>(m.size()) != equality_dim) m = zeros(equality_dim);
    } else {
      eq_multipliers.clear();
    }
    if (inequality_dim > 0) {
      if (static_cast<int>(ineq_multipliers.size()) != T) ineq_multipliers.assign(T, zeros(inequality_dim));
      else
        for (auto& m : ineq_multipliers)
          if (static_cast<int>(m.size()) != inequality_dim) m = zeros(inequality_dim);
    } else {
      ineq_multipliers.clear();
    }
  }

  // ilqr.hpp:380-407
  double compute_merit(const OCP& problem, const StateTrajectory& X, const ControlTrajectory& U) const {
    const int T = problem.horizon_steps;
    double merit = problem.objective_function ? problem.objective_function(X, U) : problem.best_cost;
    for (int t = 0; t < T; ++t) {
      if (equality_dim > 0 && problem.equality_constraints) {
        const Vec r = problem.equality_constraints(X.col(t), U.col(t));
        merit += dot(eq_multipliers[t], r) + 0.5 * penalty_parameter * dot(r, r);
      }
      if (inequality_dim > 0 && problem.inequality_constraints) {
        const Vec r = problem.inequality_constraints(X.col(t), U.col(t));
        Vec active_slack(r.size()), weighted(r.size());
        for (std::size_t i = 0; i < r.size(); ++i) {
          const double slack = r[i] > 0.0 ? r[i] : 0.0;
          const double active = (r[i] > -inequality_activation_tolerance) ? 1.0 : 0.0;
          active_slack[i] = slack * active;
          weighted[i] = ineq_multipliers[t][i] * active;
        }
        merit += dot(weighted, active_slack);
        merit += 0.5 * penalty_parameter * dot(active_slack, active_slack);
      }
    }
    return merit;
  }

  int max_iterations;
  double tolerance;
  double max_ms;
  bool debug;
  double penalty_parameter, penalty_increase, constraint_tolerance, inequality_activation_tolerance;
  int equality_dim, inequality_dim;

  std::vector<Vec> k;
  std::vector<Mat> k_matrix, a_step, b_step;
  std::vector<Vec> eq_multipliers, ineq_multipliers;
};

}  // namespace oracle
