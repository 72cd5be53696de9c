#!/bin/bash
# GPU box: sanitizer passes on a small batch, and the larger BASELINE configs
timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_debug.py n64 10 2>&1 | tail -3
timeout 300 compute-sanitizer --tool synccheck python scripts/gpu_debug.py n64 20 2>&1 | tail -2
timeout 300 compute-sanitizer --tool memcheck python scripts/gpu_debug.py n64 3 2>&1 | tail -2
for cfg in c2 c4 c5; do
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $cfg 2>&1 | tail -1 > gpurun_out/bench_r1_$cfg.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r1_$cfg.json')); c=d['config']
print('$cfg', c['workload'], 'batch', c['batch_per_gpu'], 'G,S', c['lanes_per_instance'], c['steps_per_lane'], 'value %.3e e2e %.3e kernel_ms %.3f iters %.0f evals %.1f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], c['iters_median'], c['evals_mean']))"
done
