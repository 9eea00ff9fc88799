mkdir -p gpurun_out
python tools/dense_bench.py --check --out gpurun_out/r2c_dense_before.jsonl > gpurun_out/r2c_dense_before.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_filter -s 2 -c 1 -o gpurun_out/r2c_filter_dense python tools/dense_bench.py --cases tiled:20000:1000000:4 --reps 1 > gpurun_out/r2c_ncu_filter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_rescore -s 2 -c 1 -o gpurun_out/r2c_rescore_dense python tools/dense_bench.py --cases tiled:20000:1000000:4 --reps 1 > gpurun_out/r2c_ncu_rescore.log 2>&1
cat gpurun_out/r2c_dense_before.log
