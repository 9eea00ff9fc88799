mkdir -p gpurun_out
echo "== K5 bench"
timeout 300 python tools/k5_bench.py --out gpurun_out/r2b_k5_bench.jsonl 2>&1 | tail -40
echo "== search tests"
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "route or refine or knn_search or ragged or ties or masked or prematch" 2>&1 | tail -4
echo "== cfg2 / cfg1 bench"
timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2b_bench_cfg2.json
KNNSVC_OPTIONS=concat_cluster=0 timeout 200 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r2b_bench_cfg2_onecta.json
echo "== dense bench"
timeout 400 python tools/dense_bench.py --check --cases tiled:100000:30000:4,tiled:100000:30000:32,randn:100000:1000000:4,tiled:20000:1000000:4 --out gpurun_out/r2b_dense_search2.jsonl 2>&1 | tail -12
