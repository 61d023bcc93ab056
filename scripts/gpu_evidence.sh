# Round evidence run (one gpurun call): full GPU test suite, bench (both arms), ncu launch list of one step and
# ncu --set full captures of the dominant conv kernel and of the memory-bound kernels.
# usage: bash scripts/gpu_evidence.sh <tag>      (outputs land in gpurun_out/<tag>_*)
TAG=${1:-r01}
O=gpurun_out
set -x
python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > $O/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${TAG}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.log 2>&1; echo "bench rc=$?"; tail -c 2600 $O/${TAG}_bench.log
cp $O/bench_breakdown_n1.json $O/${TAG}_per_op.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1; tail -c 1200 $O/${TAG}_bench_ref.log
CMD="python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu-baseline"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 130 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo -s 81 -c 27 -o $O/${TAG}_halo $CMD > $O/${TAG}_ncu2.log 2>&1
echo "ncu halo rc=$?"
$CMD > $O/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:pool_kernel|chanstats|im2col|cropresize|bbox|layernorm|attention|head_kernel|gate_" -s 117 -c 40 -o $O/${TAG}_mem $CMD > $O/${TAG}_ncu3.log 2>&1
echo "ncu mem rc=$?"
ls -la $O | tail -12
