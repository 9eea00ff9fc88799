mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2p_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2p_dense.jsonl 2>&1 | tee gpurun_out/r2p_dense.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2p_bench_cfg4_n1.json 2> gpurun_out/r2p_bench_cfg4_n1.err
tail -c 1500 gpurun_out/r2p_bench_cfg4_n1.json
