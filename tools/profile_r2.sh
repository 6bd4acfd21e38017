# Round-2 profiling pass (one gpurun call; every command runs plainly first, then under ncu).
set -x
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras"
D="python tools/dense_check.py solve 1225 1024 148"
$B > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_l.log 2>&1
$B > gpurun_out/plain_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"scan_rows_kernel|solve_kernel|plan_kernel|order_kernel|finalize" -s 24 -c 8 -o gpurun_out/prof_r2_tsp50 -f $B > gpurun_out/ncu_f.log 2>&1
$D > gpurun_out/plain_d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_dense.csv $D > gpurun_out/ncu_ld.log 2>&1
$D > gpurun_out/plain_d2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dense_gram_kernel|dense_solve_kernel|dense_prep_kernel" -s 3 -c 3 -o gpurun_out/prof_r2_dense -f $D > gpurun_out/ncu_fd.log 2>&1
tail -2 gpurun_out/ncu_f.log gpurun_out/ncu_fd.log | cut -c1-200
