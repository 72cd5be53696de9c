#!/bin/bash
# scratch A/B on one box: default build vs variant libraries (NEOMPC_LIB), kernel-only lines, alternating
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  python bench.py --steps 30 --config c3 --no-cpu-baseline --sustained-s 0.3 > $O/ab_new_c3_$rep.json 2> $O/ab_new_c3.err
  for v in "$@"; do
    NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_$v.so python bench.py --steps 30 --config c3 --no-cpu-baseline --sustained-s 0.3 > $O/ab_${v}_c3_$rep.json 2> $O/ab_${v}_c3.err
  done
done
python bench.py --steps 20 --config c4 --no-cpu-baseline --sustained-s 0.3 > $O/ab_new_c4.json 2> $O/ab_new_c4.err
NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_hoist.so python bench.py --steps 20 --config c4 --no-cpu-baseline --sustained-s 0.3 > $O/ab_hoist_c4.json 2> $O/ab_hoist_c4.err
