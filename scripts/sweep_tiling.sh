#!/bin/bash
# GPU box: kernel time per (config, lanes-per-instance) at large batch
for spec in "c2 65536 1" "c2 65536 2" "c2 65536 3" "c3 65536 3" "c3 65536 4" "c3 65536 5" "c3 65536 6" "c3 65536 8" "c3 65536 10" "c4 131072 5" "c4 131072 6" "c4 131072 8" "c4 131072 10" "c4 131072 16"; do
  set -- $spec
  timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config $1 --batch $2 --lanes $3 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
    print('$1 N=%d G=%d S=%d  %.3e solves/s  kernel %.3f ms  e2e %.3e  iters %.0f evals %.1f' % (c['control_steps'], c['lanes_per_instance'], c['steps_per_lane'], d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], c['iters_median'], c['evals_mean']))
except Exception as e: print('$spec failed', e)
"
done
