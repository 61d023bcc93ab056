# gate kernels with eight images per block: parity (bit-level against the fp32 reference within the test's tolerance), step
python -m pytest tests/test_memops_gpu.py tests/test_nets_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i22_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity', d['parity'], 'launches', d['gpu_launches'])"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_breakdown_n1.json"))
for o in d["per_op"][:51]:
    if o["kind"] in ("GateDesc",): print(o["i"], o["kind"], round(o["ms"],4))
PY
