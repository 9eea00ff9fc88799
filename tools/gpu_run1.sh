mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r2a_tests.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench_cfg4.json 2> gpurun_out/r2a_bench_cfg4.err
python bench.py --workload cfg4g --queries 20000 --pool 1000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_cfg4g_small.json 2> gpurun_out/r2a_bench_cfg4g_small.err
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_cfg3.json 2> gpurun_out/r2a_bench_cfg3.err
python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2a_bench_cfg5.json 2> gpurun_out/r2a_bench_cfg5.err
tail -5 gpurun_out/r2a_tests.log
