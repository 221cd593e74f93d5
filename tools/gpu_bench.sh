# usage: bash tools/gpu_bench.sh <math> [ncu]
set -x
MATH=${1:-3xbf16}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --math $MATH --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$MATH.json
if [ "$2" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv \
      --log-file gpurun_out/launches_$MATH.csv python bench.py --steps 3 --warmup 3 --math $MATH --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
  tail -3 gpurun_out/ncu_b.log
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lookup_fwd|tc_build" -s 12 -c 3 \
      -o gpurun_out/prof_$MATH -f python bench.py --steps 1 --warmup 3 --math $MATH --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
