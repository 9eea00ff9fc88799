mkdir -p gpurun_out
python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
python tools/dense_bench.py --check --out gpurun_out/r2f_dense.jsonl 2>&1 | tee gpurun_out/r2f_dense.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python tools/dense_bench.py --cases tiled:20000:1000000:4,tiled:100000:30000:32,tiled:3000:30000:32,randn:100000:1000000:4 --reps 1 > gpurun_out/r2f.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2f_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
out=[]
for r in rows[hdr+1:]:
    name=r[4][:40]; val=float(r[-1])
    if any(k in name for k in ('knn_filter','refine','rescore')): out.append((name.split('(')[0][-22:], val/1e6))
for i in range(0,len(out),4): print([f"{n}:{v:.2f}" for n,v in out[i:i+4]])
PY
