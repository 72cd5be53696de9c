#!/bin/bash
# GPU box: rebuild solve kernels with different register caps for S=3 and time config C3 / C4
for mb in 3 4 5 6; do
  touch neo_mpc_planner2_b200/csrc/kernels.cuh
  make -C neo_mpc_planner2_b200/csrc -j16 EXTRA=-DNEOMPC_MINBLOCKS_S3=$mb > /dev/null 2>&1
  grep -A1 "solve_kernelILi4ELi3E" neo_mpc_planner2_b200/csrc/build/solve_g4.ptxas.log | grep -o "Used [0-9]* registers" | head -1
  grep -A3 "solve_kernelILi4ELi3E" neo_mpc_planner2_b200/csrc/build/solve_g4.ptxas.log | grep -o "[0-9]* bytes spill stores" | head -1
  for cfg in c3 c4; do
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config $cfg 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('minblocks $mb $cfg N=%d G=%d S=%d  %.3e solves/s  kernel %.3f ms  e2e %.3e' % (c['control_steps'], c['lanes_per_instance'], c['steps_per_lane'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))"
  done
done
