#!/bin/bash
# A tuning build that differs from the product library in the single-track lane model's kernels only:
#   tools/build_st_lane_variant.sh <name> <extra nvcc flags...>  ->  tools/_variants/libmas_b200_<name>.so
# (model_st_lane.cu recompiled with the flags, every other object taken from multi_agent_solver_b200/_obj).
# Select at run time with MAS_B200_LIB=<path>.  A/B measurements only.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=$root/multi_agent_solver_b200/_obj
mkdir -p $root/tools/_variants /tmp/variant_$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC \
  -I$root/include -I$root/multi_agent_solver_b200/csrc -Xptxas -v "$@" -c $root/multi_agent_solver_b200/csrc/model_st_lane.cu \
  -o /tmp/variant_$name/model_st_lane.o > /tmp/variant_$name/ptxas.log 2>&1
others=$(ls $obj/*.o | grep -v model_st_lane.o)
nvcc -shared -o $root/tools/_variants/libmas_b200_$name.so /tmp/variant_$name/model_st_lane.o $others -ccbin /usr/bin/g++ -ldl
echo built $name
