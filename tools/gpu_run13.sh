mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/r2o_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2o_dense.jsonl 2>&1 | tee gpurun_out/r2o_dense.log
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r2o_bench_cfg1.json 2> gpurun_out/r2o_bench_cfg1.err
python bench.py --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r2o_bench_cfg2.json 2> gpurun_out/r2o_bench_cfg2.err
python bench.py --workload cfg3 --steps 30 --warmup 5 > gpurun_out/r2o_bench_cfg3.json 2> gpurun_out/r2o_bench_cfg3.err
python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2o_bench_cfg5.json 2> gpurun_out/r2o_bench_cfg5.err
for f in cfg1 cfg2 cfg3 cfg5; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2o_bench_$f.json') if l.startswith('{')][-1]); r=d.get('roofline') or {}
print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"; done
