#!/bin/bash
O=gpurun_out
timeout 70 python bench.py --config c2 --steps 20 --no-cpu-baseline --sustained-s 0.2 > $O/ab_c2_label.json 2> $O/ab_c2_label.err
timeout 60 python bench.py --steps 5 --no-cpu-baseline --sustained-s 0.2 > $O/ab_c3_label.json 2> $O/ab_c3_label.err
