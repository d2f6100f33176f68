// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Declarations of the OCP builders that live in the reference's own example sources
// (/root/reference/examples/*.cpp).  Those files are compiled UNMODIFIED: each ex_*.cpp of this
// directory is `#define main <unique name>` + `#include "<example>.cpp"`, nothing else, so the
// builder below is the reference's code, against oracle/eigen_shim.
#pragma once
#include "multi_agent_solver/ocp.hpp"

mas::OCP create_single_track_lane_following_ocp();                                   // examples/single_track_ocp.cpp:14
mas::OCP create_single_track_circular_ocp(double initial_theta, double track_radius, // examples/multi_agent_single_track.cpp:31
                                          double target_velocity, int time_steps);
mas::OCP create_linear_lqr_ocp(int n_x, int n_u, double dt, int T);                  // examples/multi_agent_lqr.cpp:21
mas::OCP create_pendulum_swingup_ocp();                                              // examples/pendulum_swing_up.cpp:29
namespace mas {
mas::OCP create_max_altitude_rocket_ocp();                                           // examples/rocket_max_altitude.cpp:31
}
