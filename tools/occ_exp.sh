python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cfg in "256 2 112640 fp64" "160 3 75776 fp64" "128 4 56320 fp64" "256 2 112640 fp32" "160 3 75776 fp32"; do
  set -- $cfg
  echo "tsp50 cfg threads=$1 ctas=$2 smem=$3 $4"
  CAVE_SOLVE_THREADS=$1 CAVE_SOLVE_CTAS_PER_SM=$2 CAVE_SOLVE_SMEM=$3 python bench.py --workload tsp50 --precision $4 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(j['value']), [round(k['ms'],3) for k in j['kernels']], j['solver']['status_counts'])"
done
