#!/bin/bash
# One parameterised runner for everything that needs the B200 box (run under gpurun from the repo root):
#   tools/gpu.sh tests [pytest args...]      GPU test suite (default: tests -m gpu -x -q)
#   tools/gpu.sh smoke                       __graft_entry__.smoke()
#   tools/gpu.sh bench [bench.py args...]    one bench line -> gpurun_out/bench_last.json (+ summary)
#   tools/gpu.sh reference                   the reference arm
#   tools/gpu.sh launches                    ncu launch list of a short bench run -> gpurun_out/launches.csv
#   tools/gpu.sh ncu <kernel-regex> <skip> [name]   ncu --set full capture of one launch -> gpurun_out/prof_<name>.ncu-rep
#   tools/gpu.sh sanitize <tool> [pytest -k expr]   compute-sanitizer memcheck|racecheck over selected GPU tests
#   tools/gpu.sh py <script> [args...]       any python script (probes under tools/)
# Several tasks can be chained with '--':  tools/gpu.sh tests -- bench -- launches
set -u
mkdir -p gpurun_out
run_one() {
  local task="$1"; shift
  case "$task" in
    tests)
      if [ $# -eq 0 ]; then set -- tests -m gpu -x -q; fi
      timeout 1500 python -m pytest -p no:cacheprovider --timeout 300 "$@" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log ;;
    bench)
      timeout 900 python bench.py "$@" 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench_last.json
      tail -3 gpurun_out/bench.err
      python tools/bench_summary.py gpurun_out/bench_last.json ;;
    reference)
      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json
      cut -c1-600 gpurun_out/bench_reference.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 60 --csv \
        --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-rows "$@" > gpurun_out/ncu_launches.log 2>&1
      tail -2 gpurun_out/ncu_launches.log ;;
    ncu)
      local k="$1" skip="$2" name="${3:-$1}"; shift; shift; [ $# -gt 0 ] && shift
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$k" -s "$skip" -c 1 \
        -o "gpurun_out/prof_$name" -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-rows "$@" > "gpurun_out/ncu_$name.log" 2>&1
      tail -2 "gpurun_out/ncu_$name.log" ;;
    sanitize)
      local tool="$1"; shift
      local expr="${1:-values_match or autograd}"
      timeout 1200 compute-sanitizer --tool "$tool" --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensorcore.py -m gpu -x -q -k "$expr" 2>&1 | tail -30 | tee "gpurun_out/sanitize_$tool.log" ;;
    py)
      timeout 1500 python "$@" 2>&1 | tail -60 ;;
    *) echo "unknown task $task"; return 2 ;;
  esac
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
args=()
for a in "$@"; do
  if [ "$a" = "--" ]; then set -x; run_one "${args[@]}"; set +x; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && { set -x; run_one "${args[@]}"; set +x; }
ls -la gpurun_out | tail -20
