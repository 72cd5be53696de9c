#!/bin/bash
# scratch A/B: default build vs variant libraries (NEOMPC_LIB), kernel-only lines
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/ab_gpu_tests.log
for cfg in c3 c4; do
  python bench.py --steps 20 --config $cfg --no-cpu-baseline --sustained-s 0.3 > $O/ab_new_$cfg.json 2> $O/ab_new_$cfg.err
done
for v in "$@"; do
  NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_$v.so python bench.py --steps 20 --config c3 --no-cpu-baseline --sustained-s 0.3 > $O/ab_${v}_c3.json 2> $O/ab_${v}_c3.err
done
