set -x
timeout 180 python -m pytest tests/test_gpu_tensorcore.py -m gpu -x -q 2>&1 | tail -30
echo "rc=$?"
nvidia-smi --query-gpu=name --format=csv,noheader
