for spec in "c3 65536 4" "c3 65536 5" "c3 65536 8" "c4 131072 8" "c4 131072 10" "c2 65536 1" "c2 65536 2"; do
  set -- $spec
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --config $1 --batch $2 --lanes $3 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['config']
print('$1 N=%d G=%d S=%d  %.3e solves/s  kernel %.3f ms' % (c['control_steps'], c['lanes_per_instance'], c['steps_per_lane'], d['value'], d['roofline']['kernel_ms']))"
done
