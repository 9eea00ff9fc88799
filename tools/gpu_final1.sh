mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r2z_tests.log
python tools/dense_bench.py --check --out gpurun_out/r2z_dense.jsonl > gpurun_out/r2z_dense.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r2z_bench_cfg4_n1.json 2> gpurun_out/r2z_bench_cfg4_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err
for w in cfg5 cfg1 cfg2 cfg3 cfg4g cfg4k32; do
  st=5; [ $w = cfg5 ] && st=2; [ $w = cfg3 ] && st=30; [ $w = cfg1 ] && st=20; [ $w = cfg2 ] && st=20
  python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline > gpurun_out/r2z_bench_${w}_n1.json 2> gpurun_out/r2z_bench_${w}_n1.err
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2z_launches_bench_100kx10M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r2z_ncu_bench.log 2>&1
for f in gpurun_out/r2z_bench_*_n1.json; do python -c "
import json,sys
d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); r=d.get('roofline') or {}
print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'filter', r.get('kernel_ms'), r.get('achieved'), 'e2e', round(d['e2e']['value'],1))"; done
