#!/bin/bash
# scratch A/B: default build vs variant libraries (NEOMPC_LIB), kernel-only lines, one ncu capture, racecheck
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/ab_gpu_tests.log
for cfg in c3 c4; do
  python bench.py --steps 20 --config $cfg --no-cpu-baseline --sustained-s 0.3 > $O/ab_new_$cfg.json 2> $O/ab_new_$cfg.err
done
for v in "$@"; do
  NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_$v.so python bench.py --steps 20 --config c3 --no-cpu-baseline --sustained-s 0.3 > $O/ab_${v}_c3.json 2> $O/ab_${v}_c3.err
done
timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_debug.py n64 10 5 2>&1 | tail -3 > $O/ab_racecheck.txt
timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_debug.py n64 20 10 2>&1 | tail -3 >> $O/ab_racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o $O/prof_ab_c3 \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustained-s 0.01 --config c3 > $O/ncu_ab_c3.log 2>&1
python scripts/latency_n1.py > $O/ab_latency_n1.txt 2>&1
NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_xp4.so python scripts/latency_n1.py > $O/ab_latency_n1_xp4.txt 2>&1
