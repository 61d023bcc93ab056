# (1) HBM write / read / copy bandwidth probes with torch (context for the write-heavy kernels), (2) fused CoordAtt3 statistics A/B
python - <<'PY'
import torch
x = torch.empty(1 << 31, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.fill_(1)); print(f"write-only fill 2 GiB: {ms:.3f} ms -> {x.numel()/ms/1e9:.2f} TB/s")
ms = t(lambda: y.copy_(x)); print(f"copy 2 GiB: {ms:.3f} ms -> {2*x.numel()/ms/1e9:.2f} TB/s (read+write)")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"read-only sum 2 GiB: {ms:.3f} ms -> {x.numel()/ms/1e9:.2f} TB/s")
PY
for k in 0 224 0 224; do UG_FUSE_STATS=$k python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i3_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_FUSE_STATS=$k', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'])"; done
tail -3 gpurun_out/i3_err.log
