// facade_host_test.cpp -- host-side behaviour of the C++ facade (include/mas_b200/mas_b200.hpp) that needs no device:
// the error conventions of the reference's interface, the name registries (examples/example_utils.hpp:19-110),
// compute_offsets (multi_agent_problem.hpp:37-50; the reference's own check is tests/ocp_tests.cpp:76-154), the OCP
// description handed to the C ABI, and -- on a box without a GPU -- that the first device call fails loudly with
// std::runtime_error instead of computing anything on the host.  Built and run by tests/test_cpu_boundary.py.
#include <cstdio>
#include <memory>
#include <string>

#include "mas_b200/mas_b200.hpp"

namespace m = mas_b200;

static int failures = 0;
#define EXPECT(cond)                                                      \
  do {                                                                    \
    if (!(cond)) {                                                        \
      std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
      ++failures;                                                         \
    }                                                                     \
  } while (0)

template <class E, class F>
static bool throws(F&& f) {
  try {
    f();
  } catch (const E&) {
    return true;
  } catch (...) {
    return false;
  }
  return false;
}

// an OCP filled in by hand, the way a caller of the reference fills its struct (no device call)
static m::OCP hand_built(int model_id, int n, int mm, int T, double dt, std::size_t id) {
  m::OCP p;
  p.model_id = model_id;
  p.state_dim = n;
  p.control_dim = mm;
  p.horizon_steps = T;
  p.dt = dt;
  p.initial_state = m::State(n, 0.0);
  p.id = id;
  return p;
}

int main() {
  // --- iLQR::set_params: .at() on the three required keys (ilqr.hpp:42-44), optional keys keep their defaults (:45-54)
  {
    m::iLQR s;
    EXPECT(throws<std::out_of_range>([&] { s.set_params({{"max_iterations", 10.0}, {"tolerance", 1e-5}}); }));
    EXPECT(throws<std::out_of_range>([&] { s.set_params({}); }));
    EXPECT(!throws<std::out_of_range>([&] { s.set_params({{"max_iterations", 10.0}, {"tolerance", 1e-5}, {"max_ms", 100.0}}); }));
    m::Solver v{std::in_place_type<m::iLQR>};
    EXPECT(throws<std::out_of_range>([&] { m::set_params(v, {{"tolerance", 1e-5}, {"max_ms", 1.0}}); }));
    mas_b200_ilqr_params d;
    mas_b200_ilqr_default_params(&d);
    EXPECT(d.max_iterations == 50 && d.tolerance == 1e-6 && d.max_ms > 1e300);                    // ctor defaults, ilqr.hpp:26-38
    EXPECT(d.penalty == 10.0 && d.penalty_increase == 5.0 && d.constraint_tolerance == 1e-4);     // :47-54
    EXPECT(d.inequality_activation_tolerance == 1e-6 && d.debug == 0);
  }
  // --- registries (example_utils.hpp:19-110): case / punctuation-insensitive keys, std::invalid_argument for unknown names
  {
    using namespace m::registry;
    EXPECT(canonical_solver_name("iLQR") == "ilqr" && canonical_solver_name("i-l_q r") == "ilqr");
    EXPECT(throws<std::invalid_argument>([] { canonical_solver_name("cgd"); }));  // not on the device path
    EXPECT(throws<std::invalid_argument>([] { make_solver("nope"); }));
    EXPECT(canonical_strategy_name("Trust-Region") == "trustregion" && canonical_strategy_name("line_search") == "linesearch");
    EXPECT(canonical_strategy_name("Centralised") == "centralized" && canonical_strategy_name("sequential_nash") == "sequential");
    const m::SolverParams prm{{"max_iterations", 10.0}, {"tolerance", 1e-5}, {"max_ms", 100.0}};
    EXPECT(throws<std::invalid_argument>([&] { make_strategy("bogus", make_solver("ilqr"), prm, 3); }));
    m::Strategy st = make_strategy("trustregion", make_solver("ilqr"), prm, 3);
    EXPECT(std::holds_alternative<m::TrustRegionNashStrategy>(st));
    EXPECT(std::holds_alternative<m::CentralizedStrategy>(make_strategy("centralized", make_solver("ilqr"), prm, 3)));
    // centralized applies the parameters to its solver at once: a missing key surfaces here (example_utils.hpp:96-99)
    EXPECT(throws<std::out_of_range>([&] { make_strategy("centralized", make_solver("ilqr"), {{"tolerance", 1e-5}}, 3); }));
  }
  // --- Matrix: column-major like Eigen::MatrixXd, per problem the bytes the C ABI takes
  {
    m::Matrix a(2, 3);
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 2; ++i) a(i, j) = 10 * i + j;
    EXPECT(a.rows() == 2 && a.cols() == 3 && a.size() == 6);
    EXPECT(a.data()[0] == 0 && a.data()[1] == 10 && a.data()[2] == 1 && a.data()[5] == 12);
    EXPECT(a.col(1) == (std::vector<double>{1, 11}));
    m::Matrix b = a;
    EXPECT(b == a);
    b.setZero();
    EXPECT(!(b == a) && b == m::Matrix::Zero(2, 3) && m::Matrix::Constant(1, 2, 7.0)(0, 1) == 7.0);
  }
  // --- OCP::verify_problem (ocp.hpp:186-236: asserts in the reference, exceptions here) and desc()
  {
    m::OCP p = hand_built(MAS_B200_MODEL_SINGLE_TRACK_LANE, 4, 2, 80, 0.1, 0);
    EXPECT(p.verify_problem());
    p.input_lower_bounds = m::Control{-0.7, -1.0};
    mas_b200_ocp_desc d = p.desc();
    EXPECT(d.has_input_bounds == 0);  // clamping needs BOTH bounds (ilqr.hpp:213)
    p.input_upper_bounds = m::Control{0.7, 1.0};
    d = p.desc();
    EXPECT(d.has_input_bounds == 1 && d.input_lower[1] == -1.0 && d.input_upper[0] == 0.7);
    EXPECT(d.model_id == MAS_B200_MODEL_SINGLE_TRACK_LANE && d.state_dim == 4 && d.control_dim == 2 && d.horizon_steps == 80 && d.dt == 0.1);
    p.input_upper_bounds = m::Control{0.7};
    EXPECT(throws<std::invalid_argument>([&] { p.verify_problem(); }));
    m::OCP q = hand_built(MAS_B200_MODEL_SINGLE_TRACK_LANE, 4, 2, 80, 0.1, 0);
    q.initial_state = m::State(3, 0.0);
    EXPECT(throws<std::invalid_argument>([&] { q.verify_problem(); }));
    m::OCP z;
    EXPECT(throws<std::invalid_argument>([&] { z.verify_problem(); }));
    m::OCP big = hand_built(MAS_B200_MODEL_LQR4, 4, 4, 10, 0.1, 0);
    big.model_params.assign(MAS_B200_MAX_PARAMS + 1, 0.0);
    EXPECT(throws<std::invalid_argument>([&] { big.desc(); }));
  }
  // --- compute_offsets: blocks sorted by agent id, running offsets (multi_agent_problem.hpp:37-50; tests/ocp_tests.cpp:76-154
  //     adds a 2x1 agent with id 5 and a 1x2 agent with id 2 and expects the id-2 block first)
  {
    m::MultiAgentProblem prob;
    prob.add_agent(std::make_shared<m::Agent>(5, std::make_shared<m::OCP>(hand_built(MAS_B200_MODEL_PENDULUM, 2, 1, 60, 0.05, 5))));
    prob.add_agent(std::make_shared<m::Agent>(2, std::make_shared<m::OCP>(hand_built(MAS_B200_MODEL_ROCKET, 3, 1, 50, 0.1, 2))));
    prob.add_agent(std::make_shared<m::Agent>(9, std::make_shared<m::OCP>(hand_built(MAS_B200_MODEL_LQR4, 4, 4, 10, 0.1, 9))));
    prob.compute_offsets();
    EXPECT(prob.blocks.size() == 3);
    EXPECT(prob.blocks[0].agent_id == 2 && prob.blocks[0].state_offset == 0 && prob.blocks[0].control_offset == 0 && prob.blocks[0].state_dim == 3);
    EXPECT(prob.blocks[1].agent_id == 5 && prob.blocks[1].state_offset == 3 && prob.blocks[1].control_offset == 1 && prob.blocks[1].control_dim == 1);
    EXPECT(prob.blocks[2].agent_id == 9 && prob.blocks[2].state_offset == 5 && prob.blocks[2].control_offset == 2 && prob.blocks[2].state_dim == 4);
    EXPECT(prob.agents[0]->id == 5);  // the agent list itself keeps insertion order
    prob.compute_offsets();           // idempotent
    EXPECT(prob.blocks.size() == 3 && prob.blocks[2].state_offset == 5);
  }
  // --- only the registered LQR shape exists on the device (multi_agent_lqr.cpp:108 uses 4 x 4)
  EXPECT(throws<std::invalid_argument>([] { m::examples::create_linear_lqr_ocp(3, 3, 0.1, 10); }));
  // --- no host implementation behind the facade: without a device the first call that needs one throws std::runtime_error
  {
    mas_b200_context_t ctx = nullptr;
    const bool have_device = mas_b200_context_create(0, nullptr, &ctx) == MAS_B200_OK;
    if (have_device) {
      mas_b200_context_destroy(ctx);
      std::printf("device present: loud-failure checks skipped\n");
    } else {
      EXPECT(throws<std::runtime_error>([] { m::examples::create_single_track_lane_following_ocp(); }));
      m::OCP p = hand_built(MAS_B200_MODEL_SINGLE_TRACK_LANE, 4, 2, 80, 0.1, 0);
      EXPECT(throws<std::runtime_error>([&] { p.initialize_problem(); }));
      p.best_controls = m::ControlTrajectory::Zero(2, 80);
      m::Solver s = m::registry::make_solver("ilqr");
      EXPECT(throws<std::runtime_error>([&] { m::solve(s, p); }));
      EXPECT(p.best_cost == std::numeric_limits<double>::max());  // nothing was computed
    }
  }
  if (failures) {
    std::printf("%d check(s) failed\n", failures);
    return 1;
  }
  std::printf("ALL OK\n");
  return 0;
}
