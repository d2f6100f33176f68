// ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's examples/pendulum_swing_up.cpp, compiled as it is
// (its main() renamed by the preprocessor so that the file can live in a shared library).
#define main ref_example_main_pendulum_swing_up
#include "pendulum_swing_up.cpp"
