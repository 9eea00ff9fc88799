python -m pytest tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -x -k "refine_route" 2>&1 | tail -4
python tools/dense_bench.py --check --cases tiled:20000:1000000:4,tiled:20000:1000000:32 2>&1 | tail -2
echo "--- ungrouped"; REFINE_GROUPED=0 python tools/dense_bench.py --cases tiled:20000:1000000:4,tiled:20000:1000000:32 2>&1 | tail -2
