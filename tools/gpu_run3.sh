set -x
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 20 --warmup 5 --math 3xbf16 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms_step',d['ms_per_step'])
for k in ('roofline','roofline_other'):
    r=d[k]; print(r['kernel'],'ms',r['ms_per_launch'],'achieved',r['achieved'],r['unit'],'frac',r['frac'],'share',r['share_of_step'])
print('e2e',d['e2e'])
"
