O=gpurun_out
q() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["arm"]
    print(sys.argv[1], "value %.3e e2e %.3e (full %.3e) kernel_ms %.4f iters %.0f evals %.1f G %d S %d"%(d["value"],d["e2e"]["value"],(d["e2e"].get("with_full_responses") or {}).get("value",0),d["roofline"]["kernel_ms"],c["iters_median"],c["evals_mean"],c["lanes_per_instance"],c["steps_per_lane"]))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
python bench.py --steps 20 --no-cpu-baseline --sustained-s 0.2 > $O/ab_default.json 2>&1; q $O/ab_default.json
python bench.py --steps 20 --no-cpu-baseline --sustained-s 0.2 --lanes 5 > $O/ab_l5.json 2>&1; q $O/ab_l5.json
for v in g5s2mb4 g5s2mb3; do NEOMPC_LIB=$PWD/neo_mpc_planner2_b200/libneompc_$v.so python bench.py --steps 20 --no-cpu-baseline --sustained-s 0.2 --lanes 5 > $O/ab_$v.json 2>&1; q $O/ab_$v.json; done
python -m pytest tests -m gpu -q 2>&1 | tail -5
