#!/bin/bash
O=gpurun_out
timeout 400 python bench.py > $O/r2_bench_c3.json 2> $O/r2_bench_c3.err
