O=gpurun_out
q() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d["arm"]
    print(sys.argv[1], "value %.3e kernel_ms %.4f sustained %.4f G %d S %d"%(d["value"],d["roofline"]["kernel_ms"],d["sustained"]["ms_per_step"],c["lanes_per_instance"],c["steps_per_lane"]))
except Exception as e: print(sys.argv[1], "failed", e)
PY
}
for l in 4 5 8; do python bench.py --steps 20 --no-cpu-baseline --sustained-s 0.3 --lanes $l > $O/tile_c3_$l.json 2>&1; q $O/tile_c3_$l.json; done
for l in 8 10 5 16; do python bench.py --steps 10 --no-cpu-baseline --sustained-s 0.3 --config c4 --lanes $l > $O/tile_c4_$l.json 2>&1; q $O/tile_c4_$l.json; done
for l in 1 2 3 4; do python bench.py --steps 20 --no-cpu-baseline --sustained-s 0.3 --config c2 --lanes $l > $O/tile_c2_$l.json 2>&1; q $O/tile_c2_$l.json; done
