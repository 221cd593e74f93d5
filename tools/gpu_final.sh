# one full GPU pass: smoke, tests, bench (+ rows, cpu baseline), reference arm, ncu launch list, ncu full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
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
for r in d.get('rows',[]): print(json.dumps(r)[:500])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
cut -c1-400 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 45 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_build" -s 1 -c 1 \
      -o gpurun_out/prof_build -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_full1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookup_fwd" -s 6 -c 1 \
      -o gpurun_out/prof_lookup -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
