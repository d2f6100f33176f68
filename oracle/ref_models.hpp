// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).  Pinned bit for bit to a build of the reference's own sources
// (oracle/_ref, tests/test_ref_pin.py; see dense.hpp).
//
// ref_models.hpp: the reference's example dynamics and OCP definitions, restated:
//   examples/models/single_track_model.hpp:23-82, pendulum_model.hpp:8-44, rocket_model.hpp:13-76
//   examples/single_track_ocp.cpp:14-116          (ST-lane,  config 1/3)
//   examples/multi_agent_single_track.cpp:31-72   (ST-circ,  config 2/5)
//   examples/multi_agent_lqr.cpp:21-76            (LQR,      config 4)
//   examples/pendulum_swing_up.cpp:29-117, examples/rocket_max_altitude.cpp:31-137
//
// Trig mode: the reference calls glibc's std::sin/cos/tan.  TRIG_GLIBC does the same.
// TRIG_PORTABLE swaps in include/mas_b200/portable_math.h, the bit-reproducible implementation the
// GPU kernels use, so GPU-vs-oracle differences can be separated into "libm last bit" (GLIBC vs
// PORTABLE oracle runs) and "everything else" (PORTABLE oracle vs GPU, expected bit-identical).
#pragma once
#include <cmath>

#include "mas_b200/portable_math.h"
#include "ref_core.hpp"

namespace oracle {

enum TrigMode { TRIG_GLIBC = 0, TRIG_PORTABLE = 1 };
inline int& trig_mode() {
  static int mode = TRIG_GLIBC;  // process-wide (read from OpenMP worker threads); set before solving
  return mode;
}
inline double o_sin(double x) { return trig_mode() == TRIG_PORTABLE ? mas_b200::pm::sin_(x) : std::sin(x); }
inline double o_cos(double x) { return trig_mode() == TRIG_PORTABLE ? mas_b200::pm::cos_(x) : std::cos(x); }
inline double o_tan(double x) { return trig_mode() == TRIG_PORTABLE ? mas_b200::pm::tan_(x) : std::tan(x); }

// ---- single_track_model.hpp:23-44 ---------------------------------------------------------------
inline Vec single_track_model(const State& x, const Control& u) {
  const double psi = x[2], v = x[3], delta = u[0], a = u[1];
  const double L = 2.5;
  Vec d(4);
  d[0] = v * o_cos(psi);
  d[1] = v * o_sin(psi);
  d[2] = v * o_tan(delta) / L;
  d[3] = a;
  return d;
}
// single_track_model.hpp:49-65
inline Mat single_track_state_jacobian(const State& x, const Control& u) {
  const double psi = x[2], v = x[3], delta = u[0];
  const double L = 2.5;
  Mat A(4, 4);
  A(0, 2) = -v * o_sin(psi);
  A(0, 3) = o_cos(psi);
  A(1, 2) = v * o_cos(psi);
  A(1, 3) = o_sin(psi);
  A(2, 3) = o_tan(delta) / L;
  return A;
}
// single_track_model.hpp:70-82
inline Mat single_track_control_jacobian(const State& x, const Control& u) {
  const double v = x[3], delta = u[0];
  const double L = 2.5;
  Mat B(4, 2);
  B(2, 0) = v / (L * o_cos(delta) * o_cos(delta));
  B(3, 1) = 1.0;
  return B;
}

// ---- single_track_ocp.cpp:14-116 ------------------------------------------------------------------
// Cost constants of the example (single_track_ocp.cpp:33-40); overridable for parameter sweeps.
struct LaneParams {
  double desired_velocity = 1.0, w_lane = 10.0, w_speed = 1.0, w_delta = 0.1, w_acc = 0.1;
};
inline OCP create_single_track_lane_following_ocp(const Vec* x0 = nullptr, const LaneParams& lp = {}) {
  OCP p;
  p.state_dim = 4;
  p.control_dim = 2;
  p.horizon_steps = 80;
  p.dt = 0.1;
  p.initial_state = x0 ? *x0 : Vec{0.0, 1.0, 0.0, 0.0};
  p.dynamics = single_track_model;
  const double desired_velocity = lp.desired_velocity, w_lane = lp.w_lane, w_speed = lp.w_speed, w_delta = lp.w_delta, w_acc = lp.w_acc;
  p.stage_cost = [=](const State& s, const Control& c, std::size_t) {
    const double y = s[1], vx = s[3], delta = c[0], a_cmd = c[1];
    const double lane_error = y, speed_error = (vx - desired_velocity);
    return w_lane * (lane_error * lane_error) + w_speed * (speed_error * speed_error) + w_delta * (delta * delta) + w_acc * (a_cmd * a_cmd);
  };
  p.terminal_cost = [](const State&) { return 0.0; };
  p.cost_state_gradient = [=](const StageCostFunction&, const State& s, const Control&, std::size_t) {
    Vec g = zeros(4);
    g[1] = 2.0 * w_lane * s[1];
    g[3] = 2.0 * w_speed * (s[3] - desired_velocity);
    return g;
  };
  p.cost_control_gradient = [=](const StageCostFunction&, const State&, const Control& c, std::size_t) {
    Vec g = zeros(2);
    g[0] = 2.0 * w_delta * c[0];
    g[1] = 2.0 * w_acc * c[1];
    return g;
  };
  p.cost_state_hessian = [=](const StageCostFunction&, const State&, const Control&, std::size_t) {
    Mat H(4, 4);
    H(1, 1) = 2.0 * w_lane;
    H(3, 3) = 2.0 * w_speed;
    return H;
  };
  p.cost_control_hessian = [=](const StageCostFunction&, const State&, const Control&, std::size_t) {
    Mat H(2, 2);
    H(0, 0) = 2.0 * w_delta;
    H(1, 1) = 2.0 * w_acc;
    return H;
  };
  p.dynamics_state_jacobian = [](const MotionModel&, const State& x, const Control& u) { return single_track_state_jacobian(x, u); };
  p.dynamics_control_jacobian = [](const MotionModel&, const State& x, const Control& u) { return single_track_control_jacobian(x, u); };
  p.input_lower_bounds = Vec{-0.7, -1.0};
  p.input_upper_bounds = Vec{0.7, 1.0};
  p.initialize_problem();
  return p;
}

// ---- lane following with path constraints ------------------------------------------------------------------
// NOT a reference example (none sets equality_constraints / inequality_constraints): the single-track lane problem
// plus  c(x,u) = a - k_gain (v_des - v) = 0  and  g(x,u) = v - v_max <= 0, used to exercise the augmented-Lagrangian
// part of iLQR::solve (ilqr.hpp:121-170,236-260,380-407).  Constraint Jacobians are the FD defaults that
// initialize_problem installs (ocp.hpp:137-171).
// jac_mask: which constraint Jacobians the problem installs itself (ocp.hpp:65-68) instead of the finite-difference defaults
// (ocp.hpp:137-171): 1 eq/state, 2 eq/control, 4 ineq/state, 8 ineq/control.
inline OCP create_single_track_lane_constrained_ocp(const Vec* x0, const LaneParams& lp, double v_max, double k_gain, int jac_mask = 0) {
  OCP p = create_single_track_lane_following_ocp(x0, lp);
  const double v_des = lp.desired_velocity;
  p.equality_constraints = [=](const State& s, const Control& c) { return Vec{c[1] - k_gain * (v_des - s[3])}; };
  p.inequality_constraints = [=](const State& s, const Control&) { return Vec{s[3] - v_max}; };
  auto row = [](std::initializer_list<double> v) {
    Mat J(1, static_cast<int>(v.size()));
    int c = 0;
    for (double e : v) J(0, c++) = e;
    return J;
  };
  if (jac_mask & 1) p.equality_constraints_state_jacobian = [=](const State&, const Control&) { return row({0.0, 0.0, 0.0, k_gain}); };
  if (jac_mask & 2) p.equality_constraints_control_jacobian = [=](const State&, const Control&) { return row({0.0, 1.0}); };
  if (jac_mask & 4) p.inequality_constraints_state_jacobian = [=](const State&, const Control&) { return row({0.0, 0.0, 0.0, 1.0}); };
  if (jac_mask & 8) p.inequality_constraints_control_jacobian = [=](const State&, const Control&) { return row({0.0, 0.0}); };
  p.initialize_problem();
  return p;
}

// ---- multi_agent_single_track.cpp:31-72 -------------------------------------------------------------
// The example derives x0 from (theta, R); a caller may pass x0 directly (batched / jittered runs).
inline Vec single_track_circular_x0(double initial_theta, double track_radius) {
  return Vec{track_radius * std::cos(initial_theta), track_radius * std::sin(initial_theta), 1.57 + initial_theta, 4.0};
}
inline OCP create_single_track_circular_ocp_from_x0(const Vec& x0, double track_radius, double target_velocity, int time_steps) {
  OCP p;
  p.state_dim = 4;
  p.control_dim = 2;
  p.horizon_steps = time_steps;
  p.dt = 0.5;
  p.initial_state = x0;
  p.dynamics = single_track_model;
  const double w_track = 1.0, w_speed = 1.0, w_delta = 0.001, w_acc = 0.001;
  p.stage_cost = [=](const State& s, const Control& c, std::size_t) {
    const double x = s[0], y = s[1], vx = s[3];
    const double delta = c[0], a_cmd = c[1];
    const double distance_from_track = std::abs(std::sqrt(x * x + y * y) - track_radius);
    const double speed_error = vx - target_velocity;
    return w_track * distance_from_track * distance_from_track + w_speed * speed_error * speed_error + w_delta * delta * delta +
           w_acc * a_cmd * a_cmd;
  };
  p.terminal_cost = [](const State&) { return 0.0; };
  p.input_lower_bounds = Vec{-0.5, -0.5};
  p.input_upper_bounds = Vec{0.5, 0.5};
  p.initialize_problem();
  return p;
}
inline OCP create_single_track_circular_ocp(double initial_theta, double track_radius, double target_velocity, int time_steps) {
  return create_single_track_circular_ocp_from_x0(single_track_circular_x0(initial_theta, track_radius), track_radius, target_velocity, time_steps);
}

// ---- multi_agent_lqr.cpp:21-76 (A = I, B = I(n_x,n_u), Q = R = Qf = I) --------------------------------
inline OCP create_linear_lqr_ocp(int n_x, int n_u, double dt, int T, const Vec* x0 = nullptr) {
  OCP p;
  p.state_dim = n_x;
  p.control_dim = n_u;
  p.dt = dt;
  p.horizon_steps = T;
  Vec init = zeros(n_x);
  if (n_x > 0) init[0] = 1.0;
  p.initial_state = x0 ? *x0 : init;
  Mat A = Mat::identity(n_x);
  Mat B(n_x, n_u);
  for (int i = 0; i < (n_x < n_u ? n_x : n_u); ++i) B(i, i) = 1.0;
  p.dynamics = [A, B](const State& x, const Control& u) { return add(matvec(A, x), matvec(B, u)); };
  p.dynamics_state_jacobian = [A](const MotionModel&, const State&, const Control&) { return A; };
  p.dynamics_control_jacobian = [B](const MotionModel&, const State&, const Control&) { return B; };
  Mat Q = Mat::identity(n_x), R = Mat::identity(n_u);
  Mat Qf = Q;
  const Mat Qt = add(Q, transpose(Q)), Rt = add(R, transpose(R)), Qf_sym = add(Qf, transpose(Qf));
  // (x^T Q) x  + (u^T R) u : row-vector times matrix first, then the inner product
  p.stage_cost = [Q, R](const State& x, const Control& u, std::size_t) { return dot(matTvec(Q, x), x) + dot(matTvec(R, u), u); };
  p.cost_state_gradient = [Qt](const StageCostFunction&, const State& x, const Control&, std::size_t) { return matvec(Qt, x); };
  p.cost_control_gradient = [Rt](const StageCostFunction&, const State&, const Control& u, std::size_t) { return matvec(Rt, u); };
  p.cost_state_hessian = [Qt](const StageCostFunction&, const State&, const Control&, std::size_t) { return Qt; };
  p.cost_control_hessian = [Rt](const StageCostFunction&, const State&, const Control&, std::size_t) { return Rt; };
  p.cost_cross_term = [n_x, n_u](const StageCostFunction&, const State&, const Control&, std::size_t) { return Mat(n_u, n_x); };
  p.terminal_cost = [Qf](const State& x) { return dot(matTvec(Qf, x), x); };
  p.terminal_cost_gradient = [Qf_sym](const TerminalCostFunction&, const State& x) { return matvec(Qf_sym, x); };
  p.terminal_cost_hessian = [Qf_sym](const TerminalCostFunction&, const State&) { return Qf_sym; };
  p.initialize_problem();
  return p;
}

// ---- pendulum_model.hpp:8-44, pendulum_swing_up.cpp:29-117 ---------------------------------------------
inline Vec pendulum_dynamics(const State& x, const Control& u) {
  const double g = 9.81, l = 1.0, m = 1.0, b = 0.1;
  Vec d(2);
  d[0] = x[1];
  d[1] = (g / l) * o_sin(x[0]) + u[0] / (m * l * l) - (b / (m * l * l)) * x[1];
  return d;
}
inline OCP create_pendulum_swingup_ocp(const Vec* x0 = nullptr) {
  OCP p;
  p.state_dim = 2;
  p.control_dim = 1;
  p.horizon_steps = 60;
  p.dt = 0.05;
  p.initial_state = x0 ? *x0 : Vec{M_PI - 0.05, 0.0};
  p.dynamics = pendulum_dynamics;
  const double g = 9.81, l = 1.0, m = 1.0;
  const double mgl = m * g * l, E_des = mgl;
  const double w_energy = 2.0, w_u = 0.05, w_shape = 2.0, w_omega = 0.05, wT_pos = 500.0, wT_vel = 100.0;
  const double horizon_d = static_cast<double>(p.horizon_steps);
  p.stage_cost = [=](const State& x, const Control& u, std::size_t k) {
    const double theta = x[0], omega = x[1], torque = u[0];
    const double s = static_cast<double>(k) / (horizon_d - 1.0);
    const double late = s * s;
    const double early = 1.0 - late;
    const double w_energy_k = w_energy * (0.2 + 0.8 * early);
    const double w_shape_k = w_shape * (0.2 + 0.8 * late);
    const double w_omega_k = w_omega * (0.2 + 0.8 * late);
    const double Tk = 0.5 * m * l * l * omega * omega;
    const double V = mgl * o_cos(theta);
    const double E = Tk + V;
    const double energy_error = (E - E_des) / mgl;
    const double upright_error = 1.0 - o_cos(theta);
    return w_energy_k * energy_error * energy_error + w_shape_k * upright_error + w_omega_k * omega * omega + w_u * torque * torque;
  };
  p.terminal_cost = [=](const State& x) {
    const double theta = x[0], omega = x[1];
    const double upright_error = 1.0 - o_cos(theta);
    return wT_pos * upright_error + wT_vel * omega * omega;
  };
  const double torque_max = 5.0;
  p.input_lower_bounds = Vec{-torque_max};
  p.input_upper_bounds = Vec{torque_max};
  p.initial_controls = ControlTrajectory(p.control_dim, p.horizon_steps);
  for (int k = 0; k < p.horizon_steps; ++k) {
    const double t = k * p.dt;
    p.initial_controls(0, k) = 0.2 * torque_max * std::sin(2.0 * M_PI * t);  // host-side setup: always glibc
  }
  p.initialize_problem();
  return p;
}

// ---- rocket_model.hpp:13-76, rocket_max_altitude.cpp:31-137 ---------------------------------------------
struct RocketParameters {
  double initial_mass = 1.0, gravity = 9.81, exhaust_velocity = 25.0;
};
inline Vec rocket_dynamics(const RocketParameters& prm, const State& s, const Control& c) {
  Vec d = zeros(3);
  const double mass = s[2] > 1e-6 ? s[2] : 1e-6;  // std::max(state(2), 1e-6)
  const double thrust = mass > 0 ? c[0] : 0.0;
  d[0] = s[1];
  d[1] = thrust / mass - prm.gravity;
  d[2] = -thrust / prm.exhaust_velocity;
  return d;
}
inline OCP create_max_altitude_rocket_ocp(const Vec* x0 = nullptr) {
  RocketParameters prm;
  prm.initial_mass = 1.0;
  prm.gravity = 9.81;
  prm.exhaust_velocity = 50.0;
  OCP p;
  p.state_dim = 3;
  p.control_dim = 1;
  p.horizon_steps = 50;
  p.dt = 0.1;
  p.initial_state = zeros(3);
  p.initial_state[2] = prm.initial_mass;
  if (x0) p.initial_state = *x0;
  p.dynamics = [prm](const State& s, const Control& c) { return rocket_dynamics(prm, s, c); };
  const double max_thrust = 20.0, w_thrust = 5e-3, w_terminal_altitude = 15.0, w_terminal_velocity = 2.0, desired_terminal_vel = 0.0;
  p.stage_cost = [=](const State&, const Control& c, std::size_t) {
    const double thrust = c[0];
    return 0.5 * w_thrust * thrust * thrust;
  };
  p.cost_control_gradient = [=](const StageCostFunction&, const State&, const Control& c, std::size_t) { return Vec{w_thrust * c[0]}; };
  p.cost_control_hessian = [=](const StageCostFunction&, const State&, const Control&, std::size_t) {
    Mat H(1, 1);
    H(0, 0) = w_thrust;
    return H;
  };
  p.cost_state_gradient = [](const StageCostFunction&, const State& s, const Control&, std::size_t) { return zeros(static_cast<int>(s.size())); };
  p.cost_state_hessian = [](const StageCostFunction&, const State& s, const Control&, std::size_t) {
    return Mat(static_cast<int>(s.size()), static_cast<int>(s.size()));
  };
  p.terminal_cost = [=](const State& s) {
    const double altitude = s[0];
    const double velocity_error = s[1] - desired_terminal_vel;
    return -w_terminal_altitude * altitude + 0.5 * w_terminal_velocity * velocity_error * velocity_error;
  };
  p.terminal_cost_gradient = [=](const TerminalCostFunction&, const State& s) {
    Vec g = zeros(static_cast<int>(s.size()));
    g[0] = -w_terminal_altitude;
    g[1] = w_terminal_velocity * (s[1] - desired_terminal_vel);
    return g;
  };
  p.terminal_cost_hessian = [=](const TerminalCostFunction&, const State& s) {
    Mat H(static_cast<int>(s.size()), static_cast<int>(s.size()));
    H(1, 1) = w_terminal_velocity;
    return H;
  };
  p.dynamics_state_jacobian = [prm](const MotionModel&, const State& s, const Control& c) {
    Mat A(3, 3);
    A(0, 1) = 1.0;
    const double thrust = c[0];
    const double mass = s[2] > 1e-6 ? s[2] : 1e-6;
    A(1, 2) = -thrust / (mass * mass);
    return A;
  };
  p.dynamics_control_jacobian = [prm](const MotionModel&, const State& s, const Control&) {
    Mat B(3, 1);
    const double mass = s[2] > 1e-6 ? s[2] : 1e-6;
    B(1, 0) = 1.0 / mass;
    B(2, 0) = -1.0 / prm.exhaust_velocity;
    return B;
  };
  p.input_lower_bounds = Vec{0.0};
  p.input_upper_bounds = Vec{max_thrust};
  p.initial_controls = ControlTrajectory(p.control_dim, p.horizon_steps);
  for (int k = 0; k < p.horizon_steps; ++k) p.initial_controls(0, k) = max_thrust / 2.0;
  p.initialize_problem();
  return p;
}

}  // namespace oracle
