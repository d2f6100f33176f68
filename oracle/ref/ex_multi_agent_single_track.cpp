// ORACLE -- TEST INFRASTRUCTURE ONLY.  The reference's examples/multi_agent_single_track.cpp, compiled as it is
// (its main() renamed by the preprocessor so that the file can live in a shared library).
#define main ref_example_main_multi_agent_single_track
#include "multi_agent_single_track.cpp"
