# ncu --set full source-level capture of kernels matching $1 in one pipeline pass (skip $2, count $3) -> CSV exports
PAT=$1; SKIP=${2:-0}; CNT=${3:-1}; TAG=${4:-kcap}
CMD="python bench.py --steps 1 --warmup 3 --batch 128 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:$PAT" -s $SKIP -c $CNT -o gpurun_out/$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/$TAG.raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/$TAG.source.csv 2>/dev/null
rm -f gpurun_out/$TAG.ncu-rep
