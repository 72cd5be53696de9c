O=gpurun_out
q() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["config"]
    print(sys.argv[1], "value %.3e e2e %.3e kernel_ms %.4f iters %.0f evals %.1f G %d S %d"%(d["value"],d["e2e"]["value"],d["roofline"]["kernel_ms"],c["iters_median"],c["evals_mean"],c["lanes_per_instance"],c["steps_per_lane"]))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
python bench.py --steps 20 --no-cpu-baseline > $O/ab_default.json 2>&1; q $O/ab_default.json
python bench.py --steps 20 --no-cpu-baseline --costmap-guidance 1 > $O/ab_unguided.json 2>&1; q $O/ab_unguided.json
NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_mb7.so python bench.py --steps 20 --no-cpu-baseline > $O/ab_mb7.json 2>&1; q $O/ab_mb7.json
for c in c2 c4 c5; do python bench.py --steps 10 --no-cpu-baseline --config $c > $O/ab_$c.json 2>&1; q $O/ab_$c.json; done
python -m pytest tests -m gpu -q 2>&1 | tail -30
