O=gpurun_out
for ctas in 0 4 8 16; do
  NEOMPC_NCCL_MAX_CTAS=$ctas python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29700 + ctas)) \
    bench.py --gpus 4 --config c3 --no-cpu-baseline --sustained-s 0.2 --steps 20 > $O/r2_ctas_${ctas}.json 2> $O/r2_ctas_${ctas}.err
done
