# round 2, second session, step 1: K5 cluster kernel + per-row decision routes
mkdir -p gpurun_out
echo "== K5 tests"
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "concat or edge_branches or row_table" > gpurun_out/r2b_k5_tests.log 2>&1; K5RC=$?
tail -15 gpurun_out/r2b_k5_tests.log; echo "k5 tests rc=$K5RC"
if [ $K5RC -eq 0 ]; then
  echo "== K5 bench"
  timeout 200 python tools/k5_bench.py --out gpurun_out/r2b_k5_bench.jsonl 2>&1 | tail -40
  OPTS=""
else
  OPTS="concat_cluster=0"
fi
echo "== full GPU suite (KNNSVC_OPTIONS=$OPTS)"
KNNSVC_OPTIONS=$OPTS timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_gpu_tests.log 2>&1; echo "suite rc=$?"
grep -v "^$" gpurun_out/r2b_gpu_tests.log | tail -25
echo "== dense bench"
timeout 400 python tools/dense_bench.py --check --cases tiled:20000:1000000:4,tiled:20000:1000000:32,G:20000:1000000:4,G:20000:1000000:32,tiled:100000:30000:4,tiled:100000:30000:32,randn:100000:1000000:4 --out gpurun_out/r2b_dense_search.jsonl 2>&1 | tail -12
