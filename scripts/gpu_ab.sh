# A/B on ONE box (box-to-box variation is +-4 %): bench with an environment switch off / on, alternating twice
# usage: bash scripts/gpu_ab.sh VAR [a b]  -> runs VAR=a / VAR=b / VAR=a / VAR=b (default 0 / 1)
V=$1
A=${2:-0}; B=${3:-1}
for k in $A $B $A $B; do
  env $V=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, clk', d['clocks']['sm_mhz'])"
done
