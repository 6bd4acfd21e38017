python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/stress.py 40 2>&1 | tail -12
for w in tsp50 tsp20 vrp20 sp5; do for r in uniform near; do
  echo "== $w $r"
  python bench.py --workload $w --regime $r --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value']), round(j['ms_per_step'],3), [round(k['ms'],3) for k in j['kernels']], j['solver']['status_counts'], j['solver']['iters_mean'], j['solver']['iters_max'], j['solve_launch_plan']['config'])"
done; done
