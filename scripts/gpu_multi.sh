# multi-GPU checks (gpurun --gpus N): NCCL shard-equality test, weak + strong scaling lines, 512-source config
N=${1:-2}
python -m pytest tests/test_dist_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -rs > gpurun_out/mg${N}_dist.log 2>&1; echo "dist tests rc=$?"; tail -6 gpurun_out/mg${N}_dist.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-yardstick "${@:2}" 2>gpurun_out/mg${N}_err.log | grep '^{' | tail -1; }
run 29711 > gpurun_out/mg${N}_weak.json; run 29712 --scaling strong --global-batch 512 > gpurun_out/mg${N}_strong512.json
run 29713 --source-size 512 > gpurun_out/mg${N}_weak_src512.json; run 29714 --collective all_gather > gpurun_out/mg${N}_weak_allgather.json
for f in weak strong512 weak_src512 weak_allgather; do python - <<PY
import json
d=json.load(open("gpurun_out/mg${N}_$f.json"))
print("$f", round(d["value"]), "img/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"]), "per-GPU batch", d["config"]["images_per_gpu_per_step"], "per-rank ms", [round(x,2) for x in d["per_rank_ms_per_step"]], "parity", d["parity"]["ok"], "gathered", d.get("gathered_result_checked"), d["clocks"]["reasons"])
PY
done
tail -3 gpurun_out/mg${N}_err.log
