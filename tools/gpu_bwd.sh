# backward rows: tests touching the backward, per-row timings, training-shaped timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "grad or bwd or backward or autograd or model or altcorr" 2>&1 | tail -5 | tee gpurun_out/pytest_bwd.log
timeout 900 python tools/bench_rows.py --reps 10 > gpurun_out/rows.jsonl 2> gpurun_out/rows.err
cut -c1-420 gpurun_out/rows.jsonl; tail -3 gpurun_out/rows.err
timeout 600 python tools/bench_bwd.py 2>&1 | tail -3 | tee gpurun_out/bench_bwd.jsonl
