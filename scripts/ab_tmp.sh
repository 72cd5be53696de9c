O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_gpu_tests_e.log
python bench.py --steps 20 > $O/r2_bench_e.json 2> $O/r2_bench_e.err
python bench.py --steps 10 --config c4 --no-cpu-baseline --sustained-s 0.3 > $O/r2_bench_e_c4.json 2>&1
python bench.py --steps 20 --config c2 --no-cpu-baseline --sustained-s 0.3 > $O/r2_bench_e_c2.json 2>&1
