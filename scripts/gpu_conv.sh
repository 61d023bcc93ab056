# conv unit parity + layer-shape sweep (used while iterating on the conv kernels)
timeout 600 python -m pytest tests/test_conv_gpu.py -m gpu -q --tb=line -p no:cacheprovider -x > gpurun_out/conv_tests.log 2>&1; echo "conv rc=$?"; tail -8 gpurun_out/conv_tests.log
timeout 600 python scripts/conv_prof.py > gpurun_out/conv_sweep.log 2>&1; cat gpurun_out/conv_sweep.log | cut -c1-200
