# tests + bench + stage probes of the lookup kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_last.json
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('value',d['value'],'ms_step',d['ms_per_step'])
for k in ('roofline','roofline_other'):
    r=d.get(k)
    if r: print(r.get('kernel'),'ms',r.get('ms_per_launch'),'bound',r['bound'],'achieved',r['achieved'],r['unit'],'frac',r['frac'],'share',r.get('share_of_step'))
print('e2e',d['e2e']); print('cpu',d.get('cpu_baseline')); print('clocks',d.get('clocks'))
for r in d.get('rows',[]): print(json.dumps(r)[:700])
PY
timeout 300 python tools/probe_bounds.py 2>&1 | tee gpurun_out/probe_bounds.jsonl
