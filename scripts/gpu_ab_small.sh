# same-box A/B of an environment switch on the latency-bound small-batch cases (UNet B=1 / B=8, GoogLeNet B=16)
# usage: bash scripts/gpu_ab_small.sh VAR [a b]
V=$1
A=${2:-0}; B=${3:-1}
for cfg in "unet 1" "unet 8" "googlenet 16"; do
  set -- $cfg
  for k in $A $B $A $B; do
    env $V=$k python bench.py --workload $1 --batch $2 --steps 200 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1 B=$2 $V=$k', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms/step')"
  done
done
