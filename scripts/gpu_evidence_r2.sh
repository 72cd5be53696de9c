#!/bin/bash
# GPU box (one B200): measurement evidence of round 2.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
nproc > $O/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_gpu_tests.log
timeout 900 python bench.py > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
for cfg in c2 c4 c5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --config $cfg > $O/r2_bench_$cfg.json 2> $O/r2_bench_$cfg.err
done
timeout 300 python bench.py --steps 20 --no-cpu-baseline --costmap-guidance 1 > $O/r2_bench_c3_unguided.json 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $O/r2_bench_ref_c3.json 2> $O/r2_bench_ref.err
# launch list of the same command as the default bench (per-launch times under ncu are cold-cache and serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-s 0.01 > $O/launches_bench_r2.log 2>&1
for cfg in c3 c2 c4 c5; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel -c 1 -o $O/prof_r2_$cfg \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustained-s 0.01 --config $cfg > $O/ncu_r2_$cfg.log 2>&1
done
timeout 300 compute-sanitizer --tool memcheck python scripts/gpu_debug.py n64 10 2>&1 | tail -2 > $O/sanitizer_r2.txt
timeout 300 compute-sanitizer --tool racecheck python scripts/gpu_debug.py n64 10 5 2>&1 | tail -2 >> $O/sanitizer_r2.txt
timeout 300 compute-sanitizer --tool synccheck python scripts/gpu_debug.py n64 20 10 2>&1 | tail -2 >> $O/sanitizer_r2.txt
neo_mpc_planner2_b200/plugin/plugin_demo latency 60 60 > $O/plugin_latency_r2.txt 2>&1
neo_mpc_planner2_b200/plugin/plugin_demo latency 1000 1000 >> $O/plugin_latency_r2.txt 2>&1
