# same-box A/B of an environment switch on one stage alone: bash scripts/gpu_ab_stage.sh VAR googlenet 256 [a b]
V=$1; W=$2; N=$3
A=${4:-0}; B=${5:-1}
for k in $A $B $A $B; do
  env $V=$k python bench.py --workload $W --batch $N --steps 30 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$W B=$N $V=$k', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms/step')"
done
