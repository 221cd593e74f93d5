set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python tools/gpu_probe.py > gpurun_out/probe.txt 2>&1; tail -20 gpurun_out/probe.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -3
