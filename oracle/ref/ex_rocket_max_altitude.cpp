// ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's examples/rocket_max_altitude.cpp, compiled as it is
// (its main() renamed by the preprocessor so that the file can live in a shared library).
#define main ref_example_main_rocket_max_altitude
#include "rocket_max_altitude.cpp"
