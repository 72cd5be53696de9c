#!/bin/bash
O=gpurun_out
for cfg in c4 c5 c2; do
timeout 60 python bench.py --config $cfg --steps 10 --no-cpu-baseline --sustained-s 0.2 > $O/ab_zc_$cfg.json 2> $O/ab_zc_$cfg.err
done
