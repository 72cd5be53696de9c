#!/bin/bash
# Tuning aid: builds libneompc_<name>.so next to libneompc.so with extra nvcc flags applied to ONE lane-group TU
# (default solve_g4.cu, the C3 kernel), reusing the other objects of the regular build.  Select it at run time with
# NEOMPC_LIB=<path> (neo_mpc_planner2_b200/_lib.py).      usage: scripts/build_variant.sh name "-DFOO=1 ..." [g4]
set -e
name=$1; extra=$2; tu=${3:-g4}
cd "$(dirname "$0")/../neo_mpc_planner2_b200/csrc"
make -j8 > /dev/null
mkdir -p build_var/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I../../include -Xptxas -v $extra \
  -c solve_$tu.cu -o build_var/$name/solve_$tu.o 2> build_var/$name/solve_$tu.ptxas.log
objs=$(ls build/*.o | grep -v "solve_$tu.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libneompc_$name.so $objs build_var/$name/solve_$tu.o
grep -A2 "solve_kernelILi4ELi3ELb0" build_var/$name/solve_$tu.ptxas.log | tail -2
