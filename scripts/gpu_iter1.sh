# iteration: K-split conv kernel — parity, per-layer timing (K-split off / on), 2-CTA MMA probe, same-box A/B of the step
python -m pytest tests/test_conv_gpu.py tests/test_stem_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i1_conv.log 2>&1; echo "conv tests rc=$?"; tail -4 gpurun_out/i1_conv.log
python -m pytest tests/test_nets_gpu.py tests/test_contract_sizes_gpu.py tests/test_memops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -rP > gpurun_out/i1_nets.log 2>&1; echo "net tests rc=$?"
grep -E "agreement|pipeline|googlenet|unet B=|argmax|passed|failed|Error|error|cls-head" gpurun_out/i1_nets.log | cut -c1-400 | head -30
SH="64,224,224,64,64,3 64,224,224,128,64,3 64,112,112,256,64,3 64,112,112,64,64,3 256,28,28,16,32,3 256,14,14,24,64,3"
for k in 0 1; do echo "== UG_KSPLIT=$k"; UG_KSPLIT=$k UG_CONFIGS=v5 timeout 300 python scripts/conv_prof.py $SH 2>&1 | tail -20; done
echo "== epilogue split (debug 8)"; UG_KSPLIT=1 UG_ABLATE=1 timeout 300 python scripts/conv_prof.py 64,224,224,64,64,3 64,224,224,128,64,3 2>&1 | tail -6
echo "== pair"; timeout 120 python scripts/mma_bench_pair.py 2>&1 | tail -14
for k in 0 1 0 1; do UG_KSPLIT=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i1_bench_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_KSPLIT=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], d['clocks']['samples'], 'parity ok', d['parity']['ok'], 'fg', round(d['parity']['mask_foreground_fraction'],4))"; done
tail -3 gpurun_out/i1_bench_err.log
