python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in tsp50 tsp20; do for prec in fp64 fp32; do
  echo "== $w $prec"
  python bench.py --workload $w --precision $prec --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value']), round(j['ms_per_step'],3), [round(k['ms'],3) for k in j['kernels']], j['solver']['status_counts'], j['solve_launch_plan']['config'])"
done; done
