set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_last.json
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('value',d['value'],'ms_step',d['ms_per_step'])
for k in ('roofline','roofline_other'):
    r=d.get(k)
    if r: print(r.get('kernel'),'ms',r.get('ms_per_launch'),'achieved',r['achieved'],r['unit'],'frac',r['frac'],'share',r.get('share_of_step'))
print('e2e',d['e2e']); print('cpu',d.get('cpu_baseline')); print('clocks',d.get('clocks'))
for r in d.get('rows',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k not in ('note',)})
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-600
