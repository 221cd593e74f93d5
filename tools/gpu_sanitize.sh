# compute-sanitizer over a subset of the GPU tests (memcheck, then racecheck on the shared-memory pipelines)
set -x
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -x -q -k "tc_volume or tc_lookup or values_match or known or far_and_nan or ondemand or autograd or bwd" 2>&1 | tail -25 | tee gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "values_match or autograd" 2>&1 | tail -25 | tee gpurun_out/sanitize_racecheck.log
