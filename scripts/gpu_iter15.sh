# attention kernel at 3 CTAs per SM (launch bounds) + the ncu part of the evidence run with the repaired kernel filter
python -m pytest tests/test_memops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "attention" 2>&1 | tail -3
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i15_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_breakdown_n1.json"))
for o in d["per_op"]:
    if o["kind"]=="AttnDesc": print(o["i"], o["kind"], round(o["ms"],4))
PY
bash scripts/gpu_evidence.sh r02d ncu 2>&1 | grep -v "^+" | tail -12
