#!/bin/bash
# scratch A/B: variant libraries (NEOMPC_LIB), kernel-only lines
O=gpurun_out
mkdir -p $O
for v in "$@"; do
  for cfg in c3 c4 c5; do
  NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_$v.so python bench.py --steps 20 --config $cfg --no-cpu-baseline --sustained-s 0.3 > $O/ab_${v}_$cfg.json 2> $O/ab_${v}_$cfg.err
  done
done
