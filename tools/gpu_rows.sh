# GPU pass: all gpu tests, per-row timings (backward, on-demand vs compiled reference), headline bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_rows.py --reps 10 > gpurun_out/rows.jsonl 2> gpurun_out/rows.err
cat gpurun_out/rows.jsonl; tail -5 gpurun_out/rows.err
timeout 600 python tools/bench_bwd.py 2>&1 | tail -3 | tee gpurun_out/bench_bwd.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_last.json
tail -3 gpurun_out/bench.err; cat gpurun_out/bench_last.json
