set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lookup_bwd|bwd_fold_pack|tc_bwd" -c 6 -o gpurun_out/prof_bwd -f python tools/bench_rows.py --reps 1 --rows bwd > gpurun_out/ncu_bwd_full.log 2>&1
tail -3 gpurun_out/ncu_bwd_full.log
ls -la gpurun_out/prof_bwd.ncu-rep
