// models.cuh -- the reference's example models as device-callable functor sets.
//
// The reference describes a problem with std::function callbacks over Eigen vectors
// (include/multi_agent_solver/types.hpp:21-50, ocp.hpp:30-81); those cannot run on a GPU, so each
// example OCP is registered here as a struct of static __host__ __device__ functions selected by
// `model_id` across the C ABI (include/mas_b200.h).  Every expression keeps the reference's
// association order, and trig goes through portable_math.h, so results are bit-reproducible
// between sm_100a and x86-64 when compiled without FMA contraction (-fmad=false).
//
//   StLane    examples/single_track_ocp.cpp:14-116 + models/single_track_model.hpp:23-82
//   StCirc    examples/multi_agent_single_track.cpp:31-72 (same dynamics, circular-track cost)
//   Lqr4      examples/multi_agent_lqr.cpp:21-76 with n_x = n_u = 4
//   Pendulum  examples/pendulum_swing_up.cpp:29-117 + models/pendulum_model.hpp:8-44
//   Rocket    examples/rocket_max_altitude.cpp:31-137 + models/rocket_model.hpp:13-76
//
// A model provides: NX, NU, NP (parameter count), AVAILABLE (which analytic derivatives exist),
// EXAMPLE_MASK (which ones the reference example installs before initialize_problem(), i.e. which
// callbacks are analytic in the reference run; the rest default to finite differences,
// ocp.hpp:117-135), dynamics / stage / terminal and the analytic derivative bodies.
#pragma once
#include "mas_b200/portable_math.h"

namespace mas_b200 {

// derivative-mode bits: set = analytic callback, clear = finite-difference default (ocp.hpp:117-135)
enum DerivBits : unsigned {
  D_A = 1u << 0,     // dynamics_state_jacobian
  D_B = 1u << 1,     // dynamics_control_jacobian
  D_LX = 1u << 2,    // cost_state_gradient
  D_LU = 1u << 3,    // cost_control_gradient
  D_LXX = 1u << 4,   // cost_state_hessian
  D_LUU = 1u << 5,   // cost_control_hessian
  D_LUX = 1u << 6,   // cost_cross_term
  D_VX = 1u << 7,    // terminal_cost_gradient
  D_VXX = 1u << 8,   // terminal_cost_hessian
  D_ALL = 0x1FFu,
  // analytic constraint Jacobians (OCP::equality_/inequality_constraints_{state,control}_jacobian, ocp.hpp:65-68); clear = the
  // central-difference default initialize_problem() installs (ocp.hpp:137-171).  Only meaningful for models with constraints.
  D_EQ_JX = 1u << 9,
  D_EQ_JU = 1u << 10,
  D_INEQ_JX = 1u << 11,
  D_INEQ_JU = 1u << 12
};

// Column-major helpers: A is NX x NX (A[r + c*NX]), B is NX x NU, l_ux is NU x NX.

// Path constraints (OCP::equality_constraints / inequality_constraints, ocp.hpp:59-66): none of the reference's
// examples sets them; models without them inherit these empty hooks and the augmented-Lagrangian code of the
// kernels (ilqr.hpp:121-170,236-260,380-407) compiles away.
//
// A_NZ / B_NZ: structural sparsity of the analytic Jacobians, bit (r + c*NX) set when entry (r, c) can be nonzero.
// The Riccati products skip the terms whose Jacobian factor is structurally zero (x + 0*y == x for finite y, so
// only the sign of an exact zero can differ).  Default: dense.
struct NoConstraints {
  static constexpr unsigned long long A_NZ = ~0ull, B_NZ = ~0ull;
  static constexpr int NEQ = 0, NINEQ = 0;
  static constexpr bool SPEC_STEP = false;  // no straight-line variant of the dynamics (rk4_step_spec falls back to rk4_step)
  MAS_HD static void eq(const double*, const double*, const double*, double*) {}
  MAS_HD static void ineq(const double*, const double*, const double*, double*) {}
  // analytic constraint Jacobians, row-major by constraint: Jx[r + c*NC] like the other column-major blocks
  MAS_HD static void eq_jac_x(const double*, const double*, const double*, double*) {}
  MAS_HD static void eq_jac_u(const double*, const double*, const double*, double*) {}
  MAS_HD static void ineq_jac_x(const double*, const double*, const double*, double*) {}
  MAS_HD static void ineq_jac_u(const double*, const double*, const double*, double*) {}
};

// ---- single-track kinematic bicycle, shared by StLane and StCirc ---------------------------------
struct SingleTrackDyn {
  // A: (0,2) (1,2) (0,3) (1,3) (2,3);  B: (2,0) (3,1)
  static constexpr unsigned long long A_NZ = (1ull << 8) | (1ull << 9) | (1ull << 12) | (1ull << 13) | (1ull << 14);
  static constexpr unsigned long long B_NZ = (1ull << 2) | (1ull << 7);
  // The control enters only through tan(delta); RK4 evaluates f four times with the same control
  // (integrator.hpp:22-25), so the tangent is computed once per step and reused: same bits, a
  // quarter of the work.
  MAS_HD static void control_terms(const double* u, double* cu) { cu[0] = pm::tan_(u[0]); }
  MAS_HD static void f_c(const double* x, const double* u, const double* cu, double* d) {
    const double psi = x[2], v = x[3], a = u[1];
    const double L = 2.5;
    double s, c;
    pm::sincos_(psi, &s, &c);
    d[0] = v * c;
    d[1] = v * s;
    d[2] = MAS_DIV_CONST(v * cu[0], L);
    d[3] = a;
  }
  MAS_HD static void f(const double* x, const double* u, double* d) {
    double cu[1];
    control_terms(u, cu);
    f_c(x, u, cu, d);
  }
  // The same two functions as straight-line code (pm::tan_spec, pm::div_const_spec): identical bits whenever *exact stays
  // true; the rollout step of the line search runs them and repeats the step with the functions above otherwise.
  MAS_HD static void control_terms_spec(const double* u, double* cu, bool* exact) { cu[0] = pm::tan_spec(u[0], exact); }
  MAS_HD static void f_c_spec(const double* x, const double* u, const double* cu, double* d, bool* exact) {
    const double psi = x[2], v = x[3], a = u[1];
    const double L = 2.5;
    double s, c;
    pm::sincos_(psi, &s, &c);
    d[0] = v * c;
    d[1] = v * s;
    d[2] = MAS_DIV_CONST_SPEC(v * cu[0], L, exact);
    d[3] = a;
  }
  MAS_HD static void jac_x(const double* x, const double* u, double* A) {
    const double psi = x[2], v = x[3], delta = u[0];
    const double L = 2.5;
    double s, c;
    pm::sincos_(psi, &s, &c);
    for (int i = 0; i < 16; ++i) A[i] = 0.0;
    A[0 + 2 * 4] = -v * s;
    A[0 + 3 * 4] = c;
    A[1 + 2 * 4] = v * c;
    A[1 + 3 * 4] = s;
    A[2 + 3 * 4] = MAS_DIV_CONST(pm::tan_(delta), L);
  }
  MAS_HD static void jac_u(const double* x, const double* u, double* B) {
    const double v = x[3], delta = u[0];
    const double L = 2.5;
    const double cd = pm::cos_(delta);
    for (int i = 0; i < 8; ++i) B[i] = 0.0;
    B[2 + 0 * 4] = v / (L * cd * cd);
    B[3 + 1 * 4] = 1.0;
  }
};

struct StLane : NoConstraints {
  static constexpr unsigned long long A_NZ = SingleTrackDyn::A_NZ, B_NZ = SingleTrackDyn::B_NZ;
  static constexpr int ID = 0;
  static constexpr int NX = 4, NU = 2, NP = 5;
  static constexpr unsigned AVAILABLE = D_A | D_B | D_LX | D_LU | D_LXX | D_LUU;
  static constexpr unsigned EXAMPLE_MASK = AVAILABLE;  // l_ux and terminal derivatives are FD
  // p = {desired_velocity, w_lane, w_speed, w_delta, w_acc}
  static constexpr int NCU = 1;
  MAS_HD static void control_terms(const double* u, const double*, double* cu) { SingleTrackDyn::control_terms(u, cu); }
  MAS_HD static void dynamics_c(const double* x, const double* u, const double* cu, const double*, double* d) { SingleTrackDyn::f_c(x, u, cu, d); }
  static constexpr bool SPEC_STEP = true;  // straight-line rollout step available (rk4_step_spec)
  MAS_HD static void control_terms_spec(const double* u, const double*, double* cu, bool* exact) { SingleTrackDyn::control_terms_spec(u, cu, exact); }
  MAS_HD static void dynamics_c_spec(const double* x, const double* u, const double* cu, const double*, double* d, bool* exact) {
    SingleTrackDyn::f_c_spec(x, u, cu, d, exact);
  }
  MAS_HD static void dynamics(const double* x, const double* u, const double*, double* d) { SingleTrackDyn::f(x, u, d); }
  MAS_HD static double stage(const double* x, const double* u, int, const double* p) {
    const double lane_error = x[1], speed_error = (x[3] - p[0]);
    const double delta = u[0], a_cmd = u[1];
    return p[1] * (lane_error * lane_error) + p[2] * (speed_error * speed_error) + p[3] * (delta * delta) + p[4] * (a_cmd * a_cmd);
  }
  MAS_HD static double terminal(const double*, const double*) { return 0.0; }
  MAS_HD static void jac_x(const double* x, const double* u, const double*, double* A) { SingleTrackDyn::jac_x(x, u, A); }
  MAS_HD static void jac_u(const double* x, const double* u, const double*, double* B) { SingleTrackDyn::jac_u(x, u, B); }
  MAS_HD static void l_x(const double* x, const double*, int, const double* p, double* g) {
    g[0] = 0.0;
    g[1] = 2.0 * p[1] * x[1];
    g[2] = 0.0;
    g[3] = 2.0 * p[2] * (x[3] - p[0]);
  }
  MAS_HD static void l_u(const double*, const double* u, int, const double* p, double* g) {
    g[0] = 2.0 * p[3] * u[0];
    g[1] = 2.0 * p[4] * u[1];
  }
  MAS_HD static void l_xx(const double*, const double*, int, const double* p, double* H) {
    for (int i = 0; i < 16; ++i) H[i] = 0.0;
    H[1 + 1 * 4] = 2.0 * p[1];
    H[3 + 3 * 4] = 2.0 * p[2];
  }
  MAS_HD static void l_uu(const double*, const double*, int, const double* p, double* H) {
    H[0] = 2.0 * p[3];
    H[1] = 0.0;
    H[2] = 0.0;
    H[3] = 2.0 * p[4];
  }
  MAS_HD static void l_ux(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void v_x(const double*, const double*, double*) {}
  MAS_HD static void v_xx(const double*, const double*, double*) {}
};

// Lane following with path constraints -- NOT one of the reference's examples (none of them has constraints):
// the registered vehicle for the augmented-Lagrangian part of iLQR::solve.  Same dynamics, cost, bounds and
// derivative mode as StLane, plus
//   equality   c(x,u) = a - k_gain * (v_des - v) = 0      (acceleration follows a proportional speed law)
//   inequality g(x,u) = v - v_max <= 0                     (speed limit)
// with finite-difference constraint Jacobians (the defaults of ocp.hpp:137-171).
// p = {desired_velocity, w_lane, w_speed, w_delta, w_acc, v_max, k_gain}
struct StLaneCon : StLane {
  static constexpr int ID = 5;
  static constexpr int NP = 7;
  static constexpr int NEQ = 1, NINEQ = 1;
  static constexpr unsigned AVAILABLE = StLane::AVAILABLE | D_EQ_JX | D_EQ_JU | D_INEQ_JX | D_INEQ_JU;
  MAS_HD static void eq(const double* x, const double* u, const double* p, double* c) { c[0] = u[1] - p[6] * (p[0] - x[3]); }
  MAS_HD static void ineq(const double* x, const double*, const double* p, double* g) { g[0] = x[3] - p[5]; }
  // d eq / d x = (0, 0, 0, k), d eq / d u = (0, 1); d ineq / d x = (0, 0, 0, 1), d ineq / d u = (0, 0)   (1 x n blocks, entry (0, c) at c)
  MAS_HD static void eq_jac_x(const double*, const double*, const double* p, double* J) {
    J[0] = 0.0;
    J[1] = 0.0;
    J[2] = 0.0;
    J[3] = p[6];
  }
  MAS_HD static void eq_jac_u(const double*, const double*, const double*, double* J) {
    J[0] = 0.0;
    J[1] = 1.0;
  }
  MAS_HD static void ineq_jac_x(const double*, const double*, const double*, double* J) {
    J[0] = 0.0;
    J[1] = 0.0;
    J[2] = 0.0;
    J[3] = 1.0;
  }
  MAS_HD static void ineq_jac_u(const double*, const double*, const double*, double* J) {
    J[0] = 0.0;
    J[1] = 0.0;
  }
};

struct StCirc : NoConstraints {
  static constexpr unsigned long long A_NZ = SingleTrackDyn::A_NZ, B_NZ = SingleTrackDyn::B_NZ;
  static constexpr int ID = 1;
  static constexpr int NX = 4, NU = 2, NP = 6;
  static constexpr unsigned AVAILABLE = D_A | D_B;
  static constexpr unsigned EXAMPLE_MASK = 0;  // the example installs no derivative callback: all FD
  // p = {track_radius, target_velocity, w_track, w_speed, w_delta, w_acc}
  static constexpr int NCU = 1;
  MAS_HD static void control_terms(const double* u, const double*, double* cu) { SingleTrackDyn::control_terms(u, cu); }
  MAS_HD static void dynamics_c(const double* x, const double* u, const double* cu, const double*, double* d) { SingleTrackDyn::f_c(x, u, cu, d); }
  static constexpr bool SPEC_STEP = true;  // straight-line rollout step available (rk4_step_spec)
  MAS_HD static void control_terms_spec(const double* u, const double*, double* cu, bool* exact) { SingleTrackDyn::control_terms_spec(u, cu, exact); }
  MAS_HD static void dynamics_c_spec(const double* x, const double* u, const double* cu, const double*, double* d, bool* exact) {
    SingleTrackDyn::f_c_spec(x, u, cu, d, exact);
  }
  MAS_HD static void dynamics(const double* x, const double* u, const double*, double* d) { SingleTrackDyn::f(x, u, d); }
  MAS_HD static double stage(const double* s, const double* c, int, const double* p) {
    const double x = s[0], y = s[1], vx = s[3];
    const double delta = c[0], a_cmd = c[1];
    const double distance_from_track = fabs(sqrt(x * x + y * y) - p[0]);
    const double speed_error = vx - p[1];
    return p[2] * distance_from_track * distance_from_track + p[3] * speed_error * speed_error + p[4] * delta * delta + p[5] * a_cmd * a_cmd;
  }
  MAS_HD static double terminal(const double*, const double*) { return 0.0; }
  MAS_HD static void jac_x(const double* x, const double* u, const double*, double* A) { SingleTrackDyn::jac_x(x, u, A); }
  MAS_HD static void jac_u(const double* x, const double* u, const double*, double* B) { SingleTrackDyn::jac_u(x, u, B); }
  MAS_HD static void l_x(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_u(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_xx(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_uu(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_ux(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void v_x(const double*, const double*, double*) {}
  MAS_HD static void v_xx(const double*, const double*, double*) {}
};

// create_linear_lqr_ocp(4, 4, dt, T): A = I, B = I, Q = R = Qf = I.  The reference evaluates the dense
// products; with identity matrices every skipped term is an exact +0, so the shortcuts below return
// the same bits for finite inputs (x^T Q x = sum of squares accumulated left to right from 0.0).
struct Lqr4 : NoConstraints {
  static constexpr unsigned long long A_NZ = 0x8421ull, B_NZ = 0x8421ull;  // identity matrices
  static constexpr int ID = 2;
  static constexpr int NX = 4, NU = 4, NP = 0;
  static constexpr unsigned AVAILABLE = D_ALL;
  static constexpr unsigned EXAMPLE_MASK = D_ALL;
  static constexpr int NCU = 1;
  MAS_HD static void control_terms(const double*, const double*, double* cu) { cu[0] = 0.0; }
  MAS_HD static void dynamics_c(const double* x, const double* u, const double*, const double* p, double* d) { dynamics(x, u, p, d); }
  MAS_HD static void dynamics(const double* x, const double* u, const double*, double* d) {
    for (int i = 0; i < 4; ++i) d[i] = x[i] + u[i];
  }
  MAS_HD static double stage(const double* x, const double* u, int, const double*) {
    double qx = x[0] * x[0];
    for (int i = 1; i < 4; ++i) qx = qx + x[i] * x[i];
    double ru = u[0] * u[0];
    for (int i = 1; i < 4; ++i) ru = ru + u[i] * u[i];
    return qx + ru;
  }
  MAS_HD static double terminal(const double* x, const double*) {
    double qx = x[0] * x[0];
    for (int i = 1; i < 4; ++i) qx = qx + x[i] * x[i];
    return qx;
  }
  MAS_HD static void jac_x(const double*, const double*, const double*, double* A) {
    for (int i = 0; i < 16; ++i) A[i] = 0.0;
    for (int i = 0; i < 4; ++i) A[i + i * 4] = 1.0;
  }
  MAS_HD static void jac_u(const double*, const double*, const double*, double* B) {
    for (int i = 0; i < 16; ++i) B[i] = 0.0;
    for (int i = 0; i < 4; ++i) B[i + i * 4] = 1.0;
  }
  MAS_HD static void l_x(const double* x, const double*, int, const double*, double* g) {
    for (int i = 0; i < 4; ++i) g[i] = 2.0 * x[i];  // (Q + Q^T) x
  }
  MAS_HD static void l_u(const double*, const double* u, int, const double*, double* g) {
    for (int i = 0; i < 4; ++i) g[i] = 2.0 * u[i];
  }
  MAS_HD static void l_xx(const double*, const double*, int, const double*, double* H) {
    for (int i = 0; i < 16; ++i) H[i] = 0.0;
    for (int i = 0; i < 4; ++i) H[i + i * 4] = 2.0;
  }
  MAS_HD static void l_uu(const double*, const double*, int, const double*, double* H) {
    for (int i = 0; i < 16; ++i) H[i] = 0.0;
    for (int i = 0; i < 4; ++i) H[i + i * 4] = 2.0;
  }
  MAS_HD static void l_ux(const double*, const double*, int, const double*, double* H) {
    for (int i = 0; i < 16; ++i) H[i] = 0.0;
  }
  MAS_HD static void v_x(const double* x, const double*, double* g) {
    for (int i = 0; i < 4; ++i) g[i] = 2.0 * x[i];
  }
  MAS_HD static void v_xx(const double*, const double*, double* H) {
    for (int i = 0; i < 16; ++i) H[i] = 0.0;
    for (int i = 0; i < 4; ++i) H[i + i * 4] = 2.0;
  }
};

struct Pendulum : NoConstraints {
  static constexpr int ID = 3;
  static constexpr int NX = 2, NU = 1, NP = 1;
  static constexpr unsigned AVAILABLE = D_A | D_B;
  static constexpr unsigned EXAMPLE_MASK = 0;  // pendulum_swing_up.cpp installs no derivative callback
  // p = {horizon_steps as double}
  static constexpr int NCU = 1;
  MAS_HD static void control_terms(const double*, const double*, double* cu) { cu[0] = 0.0; }
  MAS_HD static void dynamics_c(const double* x, const double* u, const double*, const double* p, double* d) { dynamics(x, u, p, d); }
  MAS_HD static void dynamics(const double* x, const double* u, const double*, double* d) {
    const double g = 9.81, l = 1.0, m = 1.0, b = 0.1;
    d[0] = x[1];
    d[1] = (g / l) * pm::sin_(x[0]) + u[0] / (m * l * l) - (b / (m * l * l)) * x[1];
  }
  MAS_HD static double stage(const double* x, const double* u, int k, const double* p) {
    const double g = 9.81, l = 1.0, m = 1.0;
    const double mgl = m * g * l, E_des = mgl;
    const double w_energy = 2.0, w_u = 0.05, w_shape = 2.0, w_omega = 0.05;
    const double theta = x[0], omega = x[1], torque = u[0];
    const double s = static_cast<double>(k) / (p[0] - 1.0);
    const double late = s * s;
    const double early = 1.0 - late;
    const double w_energy_k = w_energy * (0.2 + 0.8 * early);
    const double w_shape_k = w_shape * (0.2 + 0.8 * late);
    const double w_omega_k = w_omega * (0.2 + 0.8 * late);
    const double ct = pm::cos_(theta);
    const double Tk = 0.5 * m * l * l * omega * omega;
    const double V = mgl * ct;
    const double E = Tk + V;
    const double energy_error = (E - E_des) / mgl;
    const double upright_error = 1.0 - ct;
    return w_energy_k * energy_error * energy_error + w_shape_k * upright_error + w_omega_k * omega * omega + w_u * torque * torque;
  }
  MAS_HD static double terminal(const double* x, const double*) {
    const double wT_pos = 500.0, wT_vel = 100.0;
    const double upright_error = 1.0 - pm::cos_(x[0]);
    return wT_pos * upright_error + wT_vel * x[1] * x[1];
  }
  MAS_HD static void jac_x(const double* x, const double*, const double*, double* A) {
    const double g = 9.81, l = 1.0, m = 1.0, b = 0.1;
    A[0] = 0.0;
    A[1] = (g / l) * pm::cos_(x[0]);
    A[2] = 1.0;
    A[3] = -b / (m * l * l);
  }
  MAS_HD static void jac_u(const double*, const double*, const double*, double* B) {
    const double m = 1.0, l = 1.0;
    B[0] = 0.0;
    B[1] = 1.0 / (m * l * l);
  }
  MAS_HD static void l_x(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_u(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_xx(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_uu(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void l_ux(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void v_x(const double*, const double*, double*) {}
  MAS_HD static void v_xx(const double*, const double*, double*) {}
};

struct Rocket : NoConstraints {
  static constexpr int ID = 4;
  static constexpr int NX = 3, NU = 1, NP = 6;
  static constexpr unsigned AVAILABLE = D_A | D_B | D_LX | D_LU | D_LXX | D_LUU | D_VX | D_VXX;
  static constexpr unsigned EXAMPLE_MASK = AVAILABLE;  // only l_ux is FD
  // p = {gravity, exhaust_velocity, w_thrust, w_terminal_altitude, w_terminal_velocity, desired_terminal_vel}
  static constexpr int NCU = 1;
  MAS_HD static void control_terms(const double*, const double*, double* cu) { cu[0] = 0.0; }
  MAS_HD static void dynamics_c(const double* x, const double* u, const double*, const double* p, double* d) { dynamics(x, u, p, d); }
  MAS_HD static void dynamics(const double* s, const double* c, const double* p, double* d) {
    const double mass = s[2] > 1e-6 ? s[2] : 1e-6;
    const double thrust = mass > 0 ? c[0] : 0.0;
    d[0] = s[1];
    d[1] = thrust / mass - p[0];
    d[2] = -thrust / p[1];
  }
  MAS_HD static double stage(const double*, const double* c, int, const double* p) {
    const double thrust = c[0];
    return 0.5 * p[2] * thrust * thrust;
  }
  MAS_HD static double terminal(const double* s, const double* p) {
    const double altitude = s[0];
    const double velocity_error = s[1] - p[5];
    return -p[3] * altitude + 0.5 * p[4] * velocity_error * velocity_error;
  }
  MAS_HD static void jac_x(const double* s, const double* c, const double*, double* A) {
    for (int i = 0; i < 9; ++i) A[i] = 0.0;
    A[0 + 1 * 3] = 1.0;
    const double thrust = c[0];
    const double mass = s[2] > 1e-6 ? s[2] : 1e-6;
    A[1 + 2 * 3] = -thrust / (mass * mass);
  }
  MAS_HD static void jac_u(const double* s, const double*, const double* p, double* B) {
    const double mass = s[2] > 1e-6 ? s[2] : 1e-6;
    B[0] = 0.0;
    B[1] = 1.0 / mass;
    B[2] = -1.0 / p[1];
  }
  MAS_HD static void l_x(const double*, const double*, int, const double*, double* g) {
    for (int i = 0; i < 3; ++i) g[i] = 0.0;
  }
  MAS_HD static void l_u(const double*, const double* c, int, const double* p, double* g) { g[0] = p[2] * c[0]; }
  MAS_HD static void l_xx(const double*, const double*, int, const double*, double* H) {
    for (int i = 0; i < 9; ++i) H[i] = 0.0;
  }
  MAS_HD static void l_uu(const double*, const double*, int, const double* p, double* H) { H[0] = p[2]; }
  MAS_HD static void l_ux(const double*, const double*, int, const double*, double*) {}
  MAS_HD static void v_x(const double* s, const double* p, double* g) {
    g[0] = -p[3];
    g[1] = p[4] * (s[1] - p[5]);
    g[2] = 0.0;
  }
  MAS_HD static void v_xx(const double*, const double* p, double* H) {
    for (int i = 0; i < 9; ++i) H[i] = 0.0;
    H[1 + 1 * 3] = p[4];
  }
};

}  // namespace mas_b200
