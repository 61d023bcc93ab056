set -x
python -m pytest tests -m gpu -q -x --tb=short > gpurun_out/gputests_r1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/gputests_r1.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench2.log 2>&1; tail -c 2500 gpurun_out/bench2.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 1200 gpurun_out/bench_ref.log
python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 260 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 287 -c 1 -o gpurun_out/prof_r1_conv_112 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 291 -c 1 -o gpurun_out/prof_r1_conv_28 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 319 -c 2 -o gpurun_out/prof_r1_conv_224 python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline > gpurun_out/ncu4.log 2>&1
echo "ncu done rc=$?"; ls -la gpurun_out | tail -20
