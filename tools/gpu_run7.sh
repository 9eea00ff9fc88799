mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_cfg5_launches.csv python bench.py --workload cfg5 --steps 1 --warmup 1 > gpurun_out/r2g_cfg5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_cfg2_launches.csv python bench.py --workload cfg2 --steps 1 --warmup 1 > gpurun_out/r2g_cfg2.log 2>&1
python - <<'PY'
import csv, collections
for f in ('gpurun_out/r2g_cfg5_launches.csv','gpurun_out/r2g_cfg2_launches.csv'):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    agg=collections.defaultdict(lambda:[0,0.0])
    for r in rows[hdr+1:]:
        name=r[4].split('(')[0][-40:]; agg[name][0]+=1; agg[name][1]+=float(r[-1])/1e6
    tot=sum(v[1] for v in agg.values())
    print(f, 'total ms', round(tot,2))
    for n,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]: print(f'  {n:42s} n={v[0]:5d} {v[1]:9.3f} ms {100*v[1]/tot:5.1f}%')
PY
