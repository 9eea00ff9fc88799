mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches.csv python tools/dense_bench.py --cases tiled:20000:1000000:4,G:20000:1000000:32,tiled:3000:30000:32,randn:100000:1000000:4 --reps 1 > gpurun_out/r2e.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2e_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
for r in rows[hdr+1:]:
    name=r[4][:60]; val=r[-1]
    if any(k in name for k in ('knn_','prepare')): print(name, r[-2], val)
PY
