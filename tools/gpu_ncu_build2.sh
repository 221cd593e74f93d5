set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_build|lookup_fwd" -s 2 -c 2 \
      -o gpurun_out/prof_r01h -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_r01h.log 2>&1
tail -2 gpurun_out/ncu_r01h.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 45 --csv \
      --log-file gpurun_out/launches_r01h.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-rows > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
