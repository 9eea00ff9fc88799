mkdir -p gpurun_out
CASES=tiled:100000:30000:4,tiled:100000:30000:32,G:20000:1000000:4,G:20000:1000000:32,tiled:20000:1000000:4,randn:100000:1000000:4
echo "== search tests (new)"
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider -k "route or refine or knn_search or ragged or ties or masked or prematch or cfg12 or traversal or smoke" 2>&1 | tail -3
for rep in 1 2; do
echo "== dense bench NEW ($rep)"
timeout 400 python tools/dense_bench.py --check --cases $CASES --out gpurun_out/r2b_dense_qreg_new$rep.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['case'], 'total', d['total_ms'], 'filter', d['filter_ms'], 'rest', d['rest_ms'], d.get('sample_rows_equal_exact_kernel'))"
echo "== dense bench PREV ($rep)"
KNNSVC_LIB_PATH=$PWD/knn_svc_b200/libknnsvc_b200_prev.so timeout 400 python tools/dense_bench.py --check --cases $CASES --out gpurun_out/r2b_dense_qreg_prev$rep.jsonl 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['case'], 'total', d['total_ms'], 'filter', d['filter_ms'], 'rest', d['rest_ms'], d.get('sample_rows_equal_exact_kernel'))"
done
