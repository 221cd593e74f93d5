set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/bwd_launches.csv python tools/bench_rows.py --reps 1 --rows bwd > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_bwd.log
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/bwd_launches.csv')))
hdr=None
agg={}
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    k=d['Kernel Name'][:60]; m=d['Metric Name']; v=float(d['Metric Value'].replace(',',''))
    agg.setdefault((d['ID'],k),{})[m]=v
for (i,k),m in list(agg.items())[-40:]:
    print(i,k,{a.split('.')[0][-22:]:round(b,2) for a,b in m.items()})
PY
