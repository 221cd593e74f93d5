# tensor-core build: parity tests (both level-0 store paths), then the stage probes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
FLOWCORR_L0STORE=1 timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_tc_direct.log
cd tools && timeout 200 python probe_build.py | tee ../gpurun_out/probe_build.jsonl
