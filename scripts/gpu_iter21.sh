# conv1 row halo (one activation box per tile for the four row taps) + bit-identity of the pair kernels
timeout 600 python -m pytest tests/test_stem_gpu.py tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/i21_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/i21_tests.log | cut -c1-220
for k in 0 1 0 1; do UG_ROW_HALO=$k python bench.py --workload googlenet --steps 20 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('googlenet stage UG_ROW_HALO=$k', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['parity'])"; done
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i21_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_breakdown_n1.json"))
for o in d["per_op"][102:106]: print(o["i"], o["kind"], round(o["ms"],4), o.get("shape"))
PY
