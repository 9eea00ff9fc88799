mkdir -p gpurun_out
python tools/traffic_ab.py 100000 2500000 2>&1 | tee gpurun_out/r2m_traffic_ab_time.jsonl
REPS=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:knn_filter --csv --log-file gpurun_out/r2m_traffic_ab_ncu.csv python tools/traffic_ab.py 100000 2500000 > gpurun_out/r2m_traffic_ab_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2m_traffic_ab_ncu.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
cur={}
for r in rows[hdr+1:]:
    cur.setdefault(r[0],{})[r[-3]]=float(r[-1].replace(',',''))
for k,v in cur.items(): print(k, {a:round(b/1e9,2) if 'bytes' in a else b for a,b in v.items()})
PY
