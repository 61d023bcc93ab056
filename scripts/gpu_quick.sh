# conv + stem + whole-net parity, then a short bench (used while iterating on kernels)
python -m pytest tests/test_conv_gpu.py tests/test_stem_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/conv_tests.log 2>&1; echo "conv rc=$?"; tail -15 gpurun_out/conv_tests.log
python -m pytest tests/test_nets_gpu.py tests/test_memops_gpu.py -m gpu -q -rP --tb=short -p no:cacheprovider > gpurun_out/nets_tests.log 2>&1; echo "nets rc=$?"; grep -E "agreement|pipeline:|googlenet max|passed|failed|Error" gpurun_out/nets_tests.log | head -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -c 1800 gpurun_out/bench_quick.log
