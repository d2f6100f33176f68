// ORACLE -- TEST INFRASTRUCTURE ONLY (see dense.hpp header).
//
// oracle_selftest.cpp: the known answers the reference's own unit tests hold for the layers under
// iLQR, checked against the restatement.  Same inputs and expected values as
//   tests/ocp_tests.cpp:21-54   OCPTest.InitializeProblemSetsDefaultsAndBestCost
//   tests/ocp_tests.cpp:56-74   OCPTest.UpdateInitialWithBestCopiesTrajectories
//   tests/ocp_tests.cpp:76-154  MultiAgentProblemTest.BuildGlobalProblemMergesAgents
// (the fourth reference test exercises finite_differences_gradient, a CGD-only function off this path).
// Prints one line per check and exits non-zero on the first failure.
#include <cstdio>
#include <cstdlib>

#include "ref_multi_agent.hpp"

using namespace oracle;

static int g_fail = 0;
#define CHECK(cond)                                             \
  do {                                                          \
    if (!(cond)) {                                              \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); \
      g_fail = 1;                                               \
    }                                                           \
  } while (0)

static double sum(const Vec& v) {
  double s = 0.0;
  for (double x : v) s += x;
  return s;
}

static MotionModel create_integrator() {
  return [](const State& s, const Control& c) { return add(c, scale(0.0, s)); };
}

static void test_initialize_problem() {
  OCP ocp;
  ocp.state_dim = 1;
  ocp.control_dim = 1;
  ocp.horizon_steps = 3;
  ocp.dt = 0.1;
  ocp.initial_state = zeros(1);
  ocp.dynamics = create_integrator();
  ocp.stage_cost = [](const State& x, const Control& u, std::size_t) { return dot(x, x) + dot(u, u); };
  ocp.terminal_cost = [](const State& x) { return dot(x, x); };
  ocp.initialize_problem();
  CHECK(ocp.best_states.rows == 1 && ocp.best_states.cols == 4);
  CHECK(ocp.best_controls.rows == 1 && ocp.best_controls.cols == 3);
  CHECK(ocp.best_cost == 0.0);
  CHECK(static_cast<bool>(ocp.cost_state_gradient) && static_cast<bool>(ocp.cost_control_gradient));
  const Vec gx = ocp.cost_state_gradient(ocp.stage_cost, ocp.best_states.col(0), ocp.best_controls.col(0), 0);
  const Vec gu = ocp.cost_control_gradient(ocp.stage_cost, ocp.best_states.col(0), ocp.best_controls.col(0), 0);
  CHECK(gx.size() == 1 && gu.size() == 1);
  std::printf("ok initialize_problem\n");
}

static void test_update_initial_with_best() {
  OCP ocp;
  ocp.state_dim = 2;
  ocp.control_dim = 2;
  ocp.horizon_steps = 2;
  ocp.dt = 1.0;
  ocp.initial_state = zeros(2);
  ocp.dynamics = create_integrator();
  ocp.initialize_problem();
  for (auto& v : ocp.best_controls.d) v = 1.0;
  for (auto& v : ocp.best_states.d) v = 1.0;
  ocp.update_initial_with_best();
  CHECK(ocp.initial_controls.d == ocp.best_controls.d);
  CHECK(ocp.initial_states.d == ocp.best_states.d);
  std::printf("ok update_initial_with_best\n");
}

static void test_build_global_problem() {
  auto a = std::make_shared<OCP>();
  a->state_dim = 2;
  a->control_dim = 1;
  a->horizon_steps = 2;
  a->dt = 0.5;
  a->initial_state = Vec{1.0, 1.0};
  a->dynamics = [](const State& x, const Control& u) {
    Vec d(x.size());
    for (std::size_t i = 0; i < x.size(); ++i) d[i] = x[i] + u[0];
    return d;
  };
  a->stage_cost = [](const State& x, const Control& u, std::size_t) { return sum(x) + sum(u); };
  a->terminal_cost = [](const State& x) { return 2.0 * sum(x); };
  a->input_lower_bounds = Vec{-1.0};
  a->input_upper_bounds = Vec{1.0};
  a->initialize_problem();

  auto b = std::make_shared<OCP>();
  b->state_dim = 1;
  b->control_dim = 2;
  b->horizon_steps = 2;
  b->dt = 0.5;
  b->initial_state = Vec{3.0};
  b->dynamics = [](const State& x, const Control& u) {
    Vec d(x.size());
    for (std::size_t i = 0; i < x.size(); ++i) d[i] = x[i] + 2.0 * sum(u);
    return d;
  };
  b->stage_cost = [](const State& x, const Control& u, std::size_t) { return 2.0 * sum(x) + 3.0 * sum(u); };
  b->terminal_cost = [](const State& x) { return sum(x); };
  b->input_lower_bounds = Vec{-2.0, -2.0};
  b->input_upper_bounds = Vec{2.0, 2.0};
  b->initialize_problem();

  MultiAgentProblem problem;
  problem.add_agent(std::make_shared<Agent>(2, b));
  problem.add_agent(std::make_shared<Agent>(1, a));
  problem.compute_offsets();
  CHECK(problem.blocks.size() == 2);
  CHECK(problem.blocks.front().agent_id == 1 && problem.blocks.back().agent_id == 2);
  CHECK(problem.blocks.front().state_offset == 0 && problem.blocks.front().control_offset == 0);
  CHECK(problem.blocks.back().state_offset == 2 && problem.blocks.back().control_offset == 1);

  OCP g = problem.build_global_ocp();
  CHECK(g.state_dim == 3 && g.control_dim == 3 && g.horizon_steps == 2 && g.dt == 0.5);
  CHECK(g.input_lower_bounds.has_value() && g.input_upper_bounds.has_value());
  CHECK((*g.input_lower_bounds)[0] == -1.0 && (*g.input_lower_bounds)[1] == -2.0 && (*g.input_lower_bounds)[2] == -2.0);
  CHECK((g.initial_state == Vec{1.0, 1.0, 3.0}));

  const Vec state{1.0, 2.0, 3.0};     // LinSpaced(3, 1, 3)
  const Vec control{-1.0, 0.0, 1.0};  // LinSpaced(3, -1, 1)
  const Vec d = g.dynamics(state, control);
  CHECK(d.size() == 3);
  CHECK(d[0] == state[0] + control[0]);
  CHECK(d[1] == state[1] + control[0]);
  CHECK(d[2] == state[2] + 2.0 * (control[1] + control[2]));
  const double expected_stage = ((state[0] + state[1]) + control[0]) + (2.0 * state[2] + 3.0 * (control[1] + control[2]));
  CHECK(g.stage_cost(state, control, 0) == expected_stage);
  CHECK(g.terminal_cost(state) == 2.0 * (state[0] + state[1]) + state[2]);
  std::printf("ok build_global_ocp\n");
}

int main() {
  test_initialize_problem();
  test_update_initial_with_best();
  test_build_global_problem();
  if (g_fail) return 1;
  std::printf("ALL OK\n");
  return 0;
}
