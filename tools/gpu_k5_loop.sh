mkdir -p gpurun_out
echo "== K5 tests"
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "concat or edge_branches or row_table" > gpurun_out/r2b_k5_tests.log 2>&1; K5RC=$?
tail -15 gpurun_out/r2b_k5_tests.log; echo "k5 tests rc=$K5RC"
if [ $K5RC -eq 0 ]; then
  echo "== K5 bench"
  timeout 300 python tools/k5_bench.py --out gpurun_out/r2b_k5_bench.jsonl 2>&1 | grep -v general | tail -40
  echo "== cfg2 bench"
  timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2b_bench_cfg2.json | cut -c1-200
  echo "== K5 phase profile"
  KNNSVC_NVCC_EXTRA=-DKNNSVC_K5_PROFILE python -m knn_svc_b200.build --force > /dev/null 2>&1 || echo "build failed"
  timeout 120 python tools/k5_profile.py 2>&1 | tail -12
fi
