mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_gpu_tests.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed" gpurun_out/r2b_gpu_tests.log | tail -2
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2b_bench_cfg4_n1.json 2> gpurun_out/r2b_bench_cfg4_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_cfg4_n1.json').read().strip().splitlines()[-1]); print('cfg4', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'roof', d['roofline']['achieved'], d['roofline']['frac'], d['clocks'], d['verify']['ok_on_every_rank'], 'launches', d['gpu_launches'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
for W in cfg2 cfg5; do timeout 500 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_${W}_n1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2b_bench_${W}_n1.json').read().strip().splitlines()[-1]); print('$W value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],3))"; done
