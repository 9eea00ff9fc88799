mkdir -p gpurun_out
echo "== D2H overlap probe"
timeout 300 python tools/d2h_overlap_probe.py 2>&1 | tail -8 | tee gpurun_out/r2b_d2h_probe.txt
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_gpu_tests.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed" gpurun_out/r2b_gpu_tests.log | tail -3
echo "== ncu launch list, cfg2 (library kernels only)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_|concat|weight_|gather_mix|harmonic|phase_scan|prepare_rows|f0_rerank|frame_baseline|log_f0|write_plan|merge_topk" -c 400 --csv --log-file gpurun_out/r2b_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "== memcheck"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "concat_cost_both_kernels or concat_cost_batched or concat_cost_staged_long or edge_branches or row_table or decision_route or direct_route" > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2b_memcheck.log
tail -4 gpurun_out/r2b_memcheck.log
