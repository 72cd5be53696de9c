#!/bin/bash
# 8-GPU box: scaling evidence of round 2 (bench.py under torchrun, one rank per GPU; NCCL with its own CTA choice).
O=gpurun_out
run() { # n cfg extra...
  n=$1; cfg=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
    bench.py --gpus $n --config $cfg --no-cpu-baseline --sustained-s 0.3 "$@" > $O/r2_scale_${cfg}_n$n.json 2> $O/r2_scale_${cfg}_n$n.err
}
run 8 c3 --steps 20
run 8 c4 --steps 10
run 8 c5 --steps 10
run 2 c3 --steps 20
python -m pytest tests/test_fleet_native.py -m gpu -q 2>&1 | tail -3 > $O/r2_fleet_tests_n8.log
