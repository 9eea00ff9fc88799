# round 2, second session: final single-GPU measurements (run under gpurun)
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_gpu_tests.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|candidates per row|mixed \[" gpurun_out/r2b_gpu_tests.log | tail -8
echo "== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench cfg4 (default)"
timeout 600 python bench.py > gpurun_out/r2b_bench_cfg4_n1.json 2> gpurun_out/r2b_bench_cfg4_n1.err; tail -c 1500 gpurun_out/r2b_bench_cfg4_n1.json | cut -c1-1500
for W in cfg1 cfg2 cfg3 cfg4k32; do
  echo "== bench $W"
  timeout 400 python bench.py --workload $W --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_${W}_n1.json 2>/dev/null; cut -c1-260 gpurun_out/r2b_bench_${W}_n1.json | tail -1
done
echo "== bench cfg5"
timeout 500 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5_n1.json 2>/dev/null; cut -c1-330 gpurun_out/r2b_bench_cfg5_n1.json | tail -1
echo "== ncu launch list, cfg2"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_cfg2.csv python bench.py --workload cfg2 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; echo "rc=$?"
echo "== ncu full, K5 cluster kernel"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:concat_cost_cluster -c 1 -f -o gpurun_out/r2b_k5_cluster python tools/k5_profile.py > /dev/null 2>&1; echo "rc=$?"; ls -la gpurun_out/r2b_k5_cluster.ncu-rep
echo "== sanitizers"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "concat_cost_both_kernels or concat_cost_batched or edge_branches or row_table or decision_route or direct_route" > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r2b_memcheck.log
tail -4 gpurun_out/r2b_memcheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 77 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "concat_cost_both_kernels or concat_cost_batched" > gpurun_out/r2b_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/r2b_synccheck.log
tail -4 gpurun_out/r2b_synccheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "concat_cost_both_kernels and cluster or decision_route" > gpurun_out/r2b_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r2b_racecheck.log
tail -6 gpurun_out/r2b_racecheck.log
