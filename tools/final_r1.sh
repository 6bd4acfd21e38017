set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_fp64.json 2> gpurun_out/bench_fp64.err
python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
( for w in tsp50 tsp20 vrp20 sp5; do for r in uniform near; do
  python bench.py --workload $w --regime $r --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $r', round(j['value']), round(j['ms_per_step'],3), [round(k['ms'],3) for k in j['kernels']], j['solver']['status_counts'], round(j['solver']['iters_mean'],2), j['solver']['iters_max'], 'cfg', j['solve_launch_plan']['config'])"
done; done ) > gpurun_out/workloads.txt 2>&1
python tools/sweep_check.py > gpurun_out/sweep.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"scan_rows_kernel|solve_kernel|plan_kernel|order_kernel|finalize" -s 24 -c 8 -o gpurun_out/prof_r1_final3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-150
