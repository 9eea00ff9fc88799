mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5 | tee gpurun_out/r2l_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2l_dense.jsonl 2>&1 | tee gpurun_out/r2l_dense.log
python bench.py --workload cfg3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2l_bench_cfg3.json 2> gpurun_out/r2l_bench_cfg3.err
python bench.py --queries 100000 --pool 1250000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_shard8.json 2> gpurun_out/r2l_bench_shard8.err
python - <<'PY'
import json
for f in ('gpurun_out/r2l_bench_cfg3.json','gpurun_out/r2l_bench_shard8.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
    print(f, d['value'], d['ms_per_step'], 'filter', r['kernel_ms'], 'rest', d['ms_per_step']-r['kernel_ms']*r['kernel_share_of_step']*0+0, 'e2e', d['e2e']['ms_per_step'])
PY
