mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -6 | tee gpurun_out/r2q_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2q_dense.jsonl 2>&1 | tee gpurun_out/r2q_dense.log
echo "--- ungrouped"; REFINE_GROUPED=0 python tools/dense_bench.py --cases tiled:20000:1000000:4,tiled:20000:1000000:32 2>&1 | tail -2
for m in 48 100; do echo "--- refine_min $m"; REFINE_MIN=$m python tools/dense_bench.py --check --cases G:20000:1000000:32,tiled:100000:30000:4,tiled:100000:30000:32,tiled:3000:30000:32 2>&1 | tail -4; done
