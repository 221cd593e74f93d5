set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q -k "backward" 2>&1 | tail -25 | tee gpurun_out/pytest_bwd.log
timeout 600 python tools/bench_rows.py --reps 5 --rows bwd 2>&1 | tail -8 | tee gpurun_out/rows_bwd.jsonl
