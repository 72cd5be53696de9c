#!/bin/bash
# GPU box (one B200): the measurement evidence of a round — un-profiled bench lines, the ncu launch list of the same
# command, one ncu --set full capture of the hot kernel, sanitizer passes.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
nproc > $O/nproc.txt
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
for cfg in c2 c4 c5; do
  timeout 400 python bench.py --steps 10 --warmup 3 --config $cfg > $O/bench_$cfg.json 2> $O/bench_$cfg.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --footprint-mode 1 > $O/bench_c3_moving_footprint.json 2> $O/bench_c3_mf.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_c3.json 2> $O/bench_ref.err
# launch list of the same command as the default bench (per-launch times under ncu are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o $O/prof_final \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o $O/prof_moving_footprint \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --footprint-mode 1 > $O/ncu_mf.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python scripts/gpu_debug.py n64 10 2>&1 | tail -2 > $O/sanitizer.txt
timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_debug.py n64 10 5 2>&1 | tail -2 >> $O/sanitizer.txt
timeout 300 compute-sanitizer --tool synccheck python scripts/gpu_debug.py n64 20 10 2>&1 | tail -2 >> $O/sanitizer.txt
python scripts/bench_carrot.py > $O/carrot.json 2>&1
for f in c3 c2 c4 c5 c3_moving_footprint; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1]); c=d["config"]
    print("$f value %.3e e2e %.3e kernel_ms %.3f iters %.0f evals %.1f G %d S %d frac %.2e"%(d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms"],c["iters_median"],c["evals_mean"],c["lanes_per_instance"],c["steps_per_lane"],d["roofline"]["frac"]), d.get("cpu_baseline",{}).get("value"), d.get("cost_residual",{}))
except Exception as e: print("$f failed", e)
PY
done
cat $O/sanitizer.txt
