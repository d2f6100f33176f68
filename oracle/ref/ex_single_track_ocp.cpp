// ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's examples/single_track_ocp.cpp, compiled as it is
// (its main() renamed by the preprocessor so that the file can live in a shared library).
#define main ref_example_main_single_track_ocp
#include "single_track_ocp.cpp"
