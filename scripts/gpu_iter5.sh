python -m pytest tests/test_plan.py tests/test_nets_gpu.py tests/test_contract_sizes_gpu.py tests/test_memops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i5_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/i5_tests.log
echo "== GATE epilogue with / without residual loads"
UG_ABLATE=resid UG_CONFIGS=x timeout 200 python scripts/conv_prof.py 64,224,224,64,64,3 64,112,112,128,128,3 64,28,28,512,512,3 2>&1 | cut -c1-90 | tail -6
echo "== bottleneck GEMMs"
UG_CONFIGS=ffn UG_ACT=2 timeout 200 python scripts/conv_prof.py 1,1,25088,512,2048,1 2>&1 | cut -c1-110 | tail -7
UG_CONFIGS=ffn UG_ACT=0 UG_MODE=1 timeout 200 python scripts/conv_prof.py 1,1,25088,2048,512,1 1,1,25088,1024,512,1 2>&1 | cut -c1-110 | tail -14
UG_CONFIGS=ffn UG_ACT=0 timeout 200 python scripts/conv_prof.py 1,1,25088,512,2048,1 1,1,25088,512,1024,1 2>&1 | cut -c1-110 | tail -14
echo "== overlap probe"; timeout 200 python scripts/overlap_probe.py 2>&1 | tail -3
for k in 1 2; do python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i5_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/i5_err.log
