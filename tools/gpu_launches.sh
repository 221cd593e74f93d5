mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 28 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -1 gpurun_out/ncu_b.log | cut -c1-200
