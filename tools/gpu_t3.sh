set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_rows.py --reps 5 --rows bwd 2>&1 | tail -8 | tee gpurun_out/rows_bwd.jsonl
timeout 600 python tools/bench_bwd.py 2>&1 | tail -3 | tee gpurun_out/bench_bwd.jsonl
