#!/bin/bash
# GPU box: rebuild the S=3 solve kernels for different numbers of resident 64-thread blocks per SM and time C3 / C4
for mb in 6 7 8 9; do
  touch neo_mpc_planner2_b200/csrc/kernels.cuh
  make -C neo_mpc_planner2_b200/csrc -j16 EXTRA=-DNEOMPC_MINBLOCKS_RAW_S3=$mb > /dev/null 2>&1
  grep -A2 "solve_kernelILi4ELi3ELb0E" neo_mpc_planner2_b200/csrc/build/solve_g4.ptxas.log | grep -E "spill|Used"
  for cfg in c3 c4; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --config $cfg 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('resident blocks $mb $cfg N=%d G=%d S=%d  %.3e solves/s  kernel %.4f ms' % (c['control_steps'], c['lanes_per_instance'], c['steps_per_lane'], d['value'], d['roofline']['kernel_ms']))"
  done
done
