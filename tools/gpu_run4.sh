mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -40 > gpurun_out/r2d_tests.log
tail -5 gpurun_out/r2d_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2d_dense_after.jsonl 2>&1 | tee gpurun_out/r2d_dense_after.log
