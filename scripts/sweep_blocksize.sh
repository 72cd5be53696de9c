#!/bin/bash
# GPU box: rebuild with different block sizes and time C3 / C4 / C2(65536)
for bt in 32 64 128; do
  touch neo_mpc_planner2_b200/csrc/kernels.cuh
  make -C neo_mpc_planner2_b200/csrc -j16 EXTRA=-DNEOMPC_BLOCK_THREADS=$bt > /dev/null 2>&1 || echo build failed
  for spec in "c3 0" "c4 0" "c2 65536"; do
  set -- $spec
  b=""; if [ $2 != 0 ]; then b="--batch $2"; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $1 $b 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('block $bt $1 N=%d G=%d S=%d  %.3e solves/s  kernel %.3f ms  e2e %.3e' % (c['control_steps'], c['lanes_per_instance'], c['steps_per_lane'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value']))"
  done
done
