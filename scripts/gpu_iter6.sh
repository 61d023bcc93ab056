python -m pytest tests/test_conv_gpu.py tests/test_nets_gpu.py tests/test_plan.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i6_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/i6_tests.log
SH="64,112,112,64,256,0 64,56,56,128,512,0 64,28,28,256,1024,0 64,14,14,512,2048,0"
for k in 0 1; do echo "== UG_WIDE_EPI=$k"; UG_WIDE_EPI=$k UG_CONFIGS=v5 timeout 300 python scripts/conv_prof.py $SH 2>&1 | cut -c1-75,175-420 | tail -4; done
for k in 0 1 0 1; do UG_WIDE_EPI=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i6_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_WIDE_EPI=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/i6_err.log
