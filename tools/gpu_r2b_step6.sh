mkdir -p gpurun_out
echo "== new test"
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -s -k "copy_engine or pipeline or cfg12 or matcher" 2>&1 | grep -E "small read|passed|failed|Error" | tail -6
echo "== D2H overlap probe"
timeout 300 python tools/d2h_overlap_probe.py 2>&1 | tail -8 | tee gpurun_out/r2b_d2h_probe_after.txt
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_gpu_tests.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed" gpurun_out/r2b_gpu_tests.log | tail -3
echo "== cfg5"
timeout 500 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_cfg5_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_cfg5_n1.json').read().strip().splitlines()[-1]); print('cfg5 value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],1))"
echo "== cfg2 / cfg1"
for W in cfg1 cfg2; do timeout 300 python bench.py --workload $W --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_${W}_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_${W}_n1.json').read().strip().splitlines()[-1]); print('$W value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],3))"; done
