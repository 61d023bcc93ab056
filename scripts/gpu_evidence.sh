# Round evidence run (one gpurun call): full GPU test suite, smoke, bench (both arms), ncu launch list of one step and
# ncu --set full captures of the dominant conv kernels and of the memory-bound kernels (exported to CSV on the box;
# the .ncu-rep files are dropped to stay under the 64 MiB return limit).
# usage: bash scripts/gpu_evidence.sh <tag>      (outputs land in gpurun_out/<tag>_*)
TAG=${1:-r01}
O=gpurun_out
set -x
python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > $O/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -c 2800 $O/${TAG}_bench.log
cp $O/bench_breakdown_n1.json $O/${TAG}_per_op.json
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --source-size 512 > $O/${TAG}_bench_src512.log 2>&1; tail -c 600 $O/${TAG}_bench_src512.log
python bench.py --workload unet --batch 64 --steps 20 > $O/${TAG}_bench_unet64.log 2>&1; tail -c 500 $O/${TAG}_bench_unet64.log
python bench.py --workload googlenet --steps 20 > $O/${TAG}_bench_googlenet256.log 2>&1; tail -c 500 $O/${TAG}_bench_googlenet256.log
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; tail -c 1200 $O/${TAG}_bench_ref.log
CMD="python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline"
# launches per pass at batch 128: 57 (UNet) + 15 (tail ops) ... measured from the plain run's gpu_launches
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 357 -c 119 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_multi -s 132 -c 44 -o $O/${TAG}_conv $CMD > $O/${TAG}_ncu2.log 2>&1
echo "ncu conv rc=$?"
ncu -i $O/${TAG}_conv.ncu-rep --page raw --csv > $O/${TAG}_conv.raw.csv 2>/dev/null
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none -k "regex:pool_kernel|chanstats|cropresize|bbox|layernorm|attention|head_kernel|gate_|stem_conv" -s 105 -c 35 -o $O/${TAG}_mem $CMD > $O/${TAG}_ncu3.log 2>&1
echo "ncu mem rc=$?"
ncu -i $O/${TAG}_mem.ncu-rep --page raw --csv > $O/${TAG}_mem.raw.csv 2>/dev/null
rm -f $O/${TAG}_conv.ncu-rep $O/${TAG}_mem.ncu-rep
ls -la $O | tail -14
