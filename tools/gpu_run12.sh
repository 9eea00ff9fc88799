mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4 | tee gpurun_out/r2n_tests.log
for NP in 1250000 5000000; do
  CONFIGS=default REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_filter --csv --log-file gpurun_out/r2n_traffic_$NP.csv python tools/traffic_ab.py 100000 $NP > gpurun_out/r2n_traffic_$NP.log 2>&1
done
CONFIGS=cta REPS=3 python tools/traffic_ab.py 100000 1250000 2>&1 | tee gpurun_out/r2n_cta_ab.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_refine_kernel -c 1 -o gpurun_out/r2n_refine_dense python tools/dense_bench.py --cases tiled:20000:1000000:4 --reps 1 > gpurun_out/r2n_ncu_refine.log 2>&1
python bench.py --workload cfg4g --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_cfg4g.json 2> gpurun_out/r2n_bench_cfg4g.err
python bench.py --workload cfg4k32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_cfg4k32.json 2> gpurun_out/r2n_bench_cfg4k32.err
python bench.py --workload cfg3 --steps 30 --warmup 5 > gpurun_out/r2n_bench_cfg3.json 2> gpurun_out/r2n_bench_cfg3.err
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r2n_bench_cfg1.json 2> gpurun_out/r2n_bench_cfg1.err
python bench.py --workload cfg2 --steps 20 --warmup 5 > gpurun_out/r2n_bench_cfg2.json 2> gpurun_out/r2n_bench_cfg2.err
python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/r2n_bench_cfg5.json 2> gpurun_out/r2n_bench_cfg5.err
for f in cfg4g cfg4k32 cfg3 cfg1 cfg2 cfg5; do tail -c 300 gpurun_out/r2n_bench_$f.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2n_bench_$f.json') if l.startswith('{')][-1]); r=d.get('roofline') or {}
print('$f', round(d['value'],1), round(d['ms_per_step'],3), 'filter', r.get('kernel_ms'), r.get('achieved'), 'e2e', round(d['e2e']['value'],1), d.get('search_stats'))"; done
