# full GPU test suite, headline bench (no cpu baseline), per-row timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-rows 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value',d['value'],'ms_step',d['ms_per_step'])
for k in ('roofline','roofline_other'):
    r=d.get(k)
    if r: print(r.get('kernel'),'ms',r.get('ms_per_launch'),'bound',r['bound'],'achieved',r['achieved'],r['unit'],'frac',r['frac'],'share',r.get('share_of_step'))
print('e2e',d['e2e']); print('clocks',d.get('clocks'))
PY
timeout 900 python tools/bench_rows.py --reps 10 > gpurun_out/rows.jsonl 2> gpurun_out/rows.err
cut -c1-600 gpurun_out/rows.jsonl | grep -v '"math": "fp32"'; tail -3 gpurun_out/rows.err
