python -m pytest tests/test_conv_gpu.py tests/test_stem_gpu.py tests/test_memops_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/i4_conv.log 2>&1; echo "conv tests rc=$?"; tail -4 gpurun_out/i4_conv.log
python -m pytest tests/test_nets_gpu.py tests/test_contract_sizes_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -rP > gpurun_out/i4_nets.log 2>&1; echo "net tests rc=$?"
grep -E "agreement|pipeline|googlenet|unet B=|argmax|passed|failed|Error|error|cls-head" gpurun_out/i4_nets.log | cut -c1-300 | head -30
SH="64,28,28,512,512,3 64,28,28,1024,256,3 64,28,28,256,256,3 256,28,28,128,192,3"
for k in 0 1; do echo "== UG_STRIP=$k"; UG_STRIP=$k UG_CONFIGS=v5 timeout 300 python scripts/conv_prof.py $SH 2>&1 | cut -c1-70 | tail -8; done
for v in "UG_STRIP=0 UG_FUSE_HEAD=0" "UG_STRIP=1 UG_FUSE_HEAD=0" "UG_STRIP=1 UG_FUSE_HEAD=1" "UG_STRIP=0 UG_FUSE_HEAD=0" "UG_STRIP=1 UG_FUSE_HEAD=1"; do env $v python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i4_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'])"; done
tail -3 gpurun_out/i4_err.log
python bench.py --workload googlenet --steps 20 | tail -1 | cut -c1-300
UG_FUSE_HEAD=0 python bench.py --workload googlenet --steps 20 | tail -1 | cut -c1-300
