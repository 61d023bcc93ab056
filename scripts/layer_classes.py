"""Per-op event timings of one step (bench.py's breakdown JSON) grouped by layer class: share of the step and TFLOP/s.
usage: layer_classes.py gpurun_out/bench_breakdown_n1.json > profiles/rNN_layer_classes.md"""
import json, sys
d = json.load(open(sys.argv[1]))
ops = d["per_op"]
first_g = next(i for i, o in enumerate(ops) if o["kind"] == "S2dDesc" or (o["kind"] == "StemDesc" and o["shape"][5] == 7))   # GoogLeNet conv1


def cls(i, o):
    k, sh = o["kind"], o["shape"]
    g = i >= first_g
    if k == "StemDesc":
        return "stem convs (inc, conv1)"
    if k == "ConvDesc":
        B, H, W, Cin, N, R = sh
        if g and R == 4:
            return "stem convs (inc, conv1)"        # conv1 as a four-row-tap GEMM over the space-to-depth image
        if g:
            return "GoogLeNet 3x3" if R == 3 else "GoogLeNet 1x1 (fused heads, branch4, conv2)"
        if R == 3:
            return f"UNet 3x3 {H}x{W} N={'64' if N <= 64 else '>=128'}"
        return "UNet 1x1 (ConvTranspose, linear layers)"
    return {"PoolDesc": "max-pools (GoogLeNet)", "ChanStatsDesc": "CoordAtt3 statistics + gate", "GateDesc": "CoordAtt3 statistics + gate",
            "S2dDesc": "stem convs (inc, conv1)", "LayerNormDesc": "LayerNorm + attention", "AttnDesc": "LayerNorm + attention"}.get(k, "bbox / crop-resize / head / front-end")


acc = {}
for i, o in enumerate(ops):
    c = acc.setdefault(cls(i, o), [0, 0.0, 0.0])
    c[0] += 1
    c[1] += o["ms"]
    c[2] += o["gflop"] or 0.0
tot = sum(v[1] for v in acc.values())
print(f"# Per-op event timings of one {d['images_per_step']}-image step grouped by layer class (source: the bench breakdown JSON; per-op timing adds")
print("# the launch gaps that the back-to-back step hides, so the classes sum to slightly more than ms_per_step)\n")
print("| layer class | launches | ms / step | share | TFLOP/s |\n|---|---|---|---|---|")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    tf = f"{v[2] / v[1]:.0f}" if v[2] else "-"
    print(f"| {k} | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f} % | {tf} |")
print(f"| total | {sum(v[0] for v in acc.values())} | {tot:.2f} | 100 % | {sum(v[2] for v in acc.values()) / tot:.0f} |")
