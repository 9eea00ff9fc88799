mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "refine_route or row_table or gather_mix_sharded or merge_topk64 or accumulator or knn_search_vs or ragged or ties_and or block_traversal or masked" > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2_memcheck.log
tail -5 gpurun_out/r2_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -p no:cacheprovider -k "refine_route or merge_topk64 or gather_mix_sharded" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2_racecheck.log
tail -8 gpurun_out/r2_racecheck.log
