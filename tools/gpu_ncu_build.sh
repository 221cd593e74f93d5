set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_build" -s 3 -c 1 \
      -o gpurun_out/prof_build -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 20 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
