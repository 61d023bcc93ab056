python -m pytest tests/test_stem_gpu.py -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/stem_tests.log 2>&1; echo "stem rc=$?"; tail -3 gpurun_out/stem_tests.log
CMD="python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:stem_conv -s 6 -c 2 -o gpurun_out/r01d_stem $CMD > gpurun_out/ncu_stem.log 2>&1
echo "ncu rc=$?"; tail -c 600 gpurun_out/plain.log
ncu -i gpurun_out/r01d_stem.ncu-rep --page raw --csv > gpurun_out/r01d_stem.raw.csv 2>/dev/null
ncu -i gpurun_out/r01d_stem.ncu-rep --page source --csv > gpurun_out/r01d_stem.source.csv 2>/dev/null
ls -la gpurun_out
