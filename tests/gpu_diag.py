"""Run every GPU test case in crash-isolated fashion and log one line per test to gpurun_out/diag.log.
Used during bring-up: a trapped kernel kills only its own pytest process, the loop continues with the next file."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
files = sys.argv[1:] or ["tests/test_conv_gpu.py", "tests/test_memops_gpu.py"]
with open(os.path.join(ROOT, "gpurun_out", "diag.log"), "w") as log:
    for f in files:
        r = subprocess.run([sys.executable, "-m", "pytest", f, "-m", "gpu", "-q", "-rA", "--tb=line", "-p",
                            "no:cacheprovider"], cwd=ROOT, capture_output=True, text=True, timeout=900)
        log.write(f"==== {f} rc={r.returncode}\n{r.stdout[-12000:]}\n{r.stderr[-3000:]}\n")
        log.flush()
        print(f"==== {f} rc={r.returncode}")
        print(r.stdout[-6000:])
