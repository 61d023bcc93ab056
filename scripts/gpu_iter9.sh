python -m pytest tests/test_stem_gpu.py tests/test_conv_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i9_conv.log 2>&1; echo "stem+conv tests rc=$?"; tail -8 gpurun_out/i9_conv.log
python -m pytest tests/test_nets_gpu.py tests/test_contract_sizes_gpu.py tests/test_plan.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i9_nets.log 2>&1; echo "net tests rc=$?"; tail -3 gpurun_out/i9_nets.log
for k in 0 1; do UG_CONV1_S2D=$k python bench.py --workload googlenet --steps 20 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('googlenet stage UG_CONV1_S2D=$k', round(d['value']), 'img/s', round(d['ms_per_step'],3), 'ms', d['parity'])"; done
for k in 0 1 0 1; do UG_CONV1_S2D=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i9_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_CONV1_S2D=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/i9_err.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_breakdown_n1.json"))
for o in d["per_op"][106:114]: print(o["i"], o["kind"], round(o["ms"],4), o["shape"])
PY
