set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_last.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_last.json'))
print('value',d['value'],'ms_step',d['ms_per_step'])
for k in ('roofline','roofline_other'):
    r=d[k]; print(r['kernel'],'ms',r['ms_per_launch'],'achieved',r['achieved'],r['unit'],'frac',r['frac'],'share',r['share_of_step'])
print('e2e',d['e2e'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lookup_fwd|tc_build" -s 12 -c 2 \
      -o gpurun_out/prof_last -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
