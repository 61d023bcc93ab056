# A/B on ONE box (box-to-box variation is +-4 %): bench with an environment switch off / on, alternating twice
# usage: bash scripts/gpu_ab.sh VAR   -> runs VAR=0 / VAR=1 / VAR=0 / VAR=1
V=$1
for k in 0 1 0 1; do
  env $V=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$V=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, clk', d['clocks']['sm_mhz'])"
done
