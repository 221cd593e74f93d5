set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_build" -s 1 -c 1 \
      -o gpurun_out/prof_r01h_build -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_r01h_build.log 2>&1
tail -2 gpurun_out/ncu_r01h_build.log
