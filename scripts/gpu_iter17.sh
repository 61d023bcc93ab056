# sliding 3x3 stride-2 max-pool kernel: parity, GoogLeNet stage A/B
python -m pytest tests/test_memops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "pool" 2>&1 | tail -3
for k in 0 1 0 1; do UG_POOL_BLOCK=$k python bench.py --workload googlenet --steps 20 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('googlenet stage UG_POOL_BLOCK=$k', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['parity']['ok'])"; done
UG_POOL_BLOCK=1 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i17_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_breakdown_n1.json"))
for o in d["per_op"]:
    if o["kind"]=="PoolDesc": print(o["i"], o["kind"], round(o["ms"],4))
PY
