#!/bin/bash
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_horizon" 2>&1 | tail -15 > $O/ab_fullhorizon_test.log
