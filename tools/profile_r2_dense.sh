# Round-2 dense-path capture on the final build (plain run first, then ncu).
set -x
D="python tools/dense_check.py solve 1225 1024 148"
$D > gpurun_out/plain_d3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dense_gram_kernel|dense_solve_kernel|dense_prep_kernel" -s 3 -c 3 -o gpurun_out/prof_r2_dense_final -f $D > gpurun_out/ncu_fd2.log 2>&1
tail -n 2 gpurun_out/ncu_fd2.log | cut -c1-200
