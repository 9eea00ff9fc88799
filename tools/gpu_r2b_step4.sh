mkdir -p gpurun_out
echo "== search tests"
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -s -k "grouped or route or refine or knn_search or ragged or ties or masked or prematch or cfg12 or accumulator or traversal" > gpurun_out/r2b_search_tests.log 2>&1; RC=$?
grep -E "candidates per row|passed|failed|Error" gpurun_out/r2b_search_tests.log | tail -12; echo "search tests rc=$RC"
CASES=tiled:100000:30000:4,tiled:100000:30000:32,G:20000:1000000:32,tiled:3000:30000:32
echo "== dense bench, grouped direct route"
timeout 300 python tools/dense_bench.py --check --cases $CASES --out gpurun_out/r2b_dense_group_on.jsonl 2>&1 | cut -c1-260 | tail -6
echo "== dense bench, per-row direct route"
KNNSVC_OPTIONS=rescore_group=0 timeout 300 python tools/dense_bench.py --check --cases $CASES --out gpurun_out/r2b_dense_group_off.jsonl 2>&1 | cut -c1-260 | tail -6
if [ $RC -eq 0 ]; then
echo "== cfg5"
timeout 400 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2b_bench_cfg5_n1.json | cut -c1-330
fi
