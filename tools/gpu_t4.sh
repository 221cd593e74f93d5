set -x
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k training_step 2>&1 | grep -E "Error|error|fc_|passed|failed" | head -20
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward or gradients or two_blocks" 2>&1 | grep -E "Error|error|fc_|passed|failed" | head -20
CUDA_LAUNCH_BLOCKING=1 timeout 600 python tools/bench_rows.py --reps 2 --rows bwd 2>&1 | tail -4 | cut -c1-300
