#!/bin/bash
# Builds the example programs against libmas_b200.so and links them under the reference's binary names.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
mkdir -p "$HERE/bin"
/usr/bin/g++ -std=c++17 -O2 -I"$ROOT/include" "$HERE/mas_b200_examples.cpp" -o "$HERE/bin/mas_b200_examples" \
  -L"$ROOT/multi_agent_solver_b200" -lmas_b200 -Wl,-rpath,"$ROOT/multi_agent_solver_b200"
for n in single_track_ocp pendulum_swing_up rocket_max_altitude multi_agent_single_track multi_agent_lqr multi_agent_mixed; do
  ln -sf mas_b200_examples "$HERE/bin/$n"
done
echo "built $HERE/bin"
