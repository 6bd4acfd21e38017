run() { python bench.py --workload $1 --batch $2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value']), round(j['ms_per_step'],3), [round(k['ms'],3) for k in j['kernels']], j['solver']['status_counts'], j['solver']['loss'])"; }
for o in 0 1 1; do echo "== order $o tsp50 4096"; CAVE_SOLVE_ORDER=$o run tsp50 4096; done
echo "== 16384"; run tsp50 16384
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
grep -E "order_kernel|plan_kernel" gpurun_out/launches_r1c.csv | tail -4 | cut -c1-60,150-260
