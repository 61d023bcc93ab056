# CTA-pair kernel, second step: GATE epilogue (residual by TMA) and one-stream mode for 192 / 256 input channels
timeout 300 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "variant6" > gpurun_out/i11_pair.log 2>&1; echo "pair tests rc=$?"; tail -15 gpurun_out/i11_pair.log
UG_CONFIGS=pair timeout 300 python scripts/conv_prof.py 64,224,224,64,64,3 64,112,112,256,64,3 64,112,112,64,64,3 > gpurun_out/i11_prof.log 2>&1; cat gpurun_out/i11_prof.log | cut -c1-110 | tail -20
for k in 0 1 0 1; do UG_PAIR=$k timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i11_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_PAIR=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/i11_err.log
