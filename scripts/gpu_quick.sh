# whole GPU suite + smoke + a short bench (used while iterating); outputs in gpurun_out/q_*
python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider -rP > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"
grep -E "agreement|pipeline|googlenet|unet B=|passed|failed|Error|error|cls-head" gpurun_out/q_tests.log | head -40; tail -5 gpurun_out/q_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/q_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/q_bench.log 2>&1; echo "bench rc=$?"; tail -c 4500 gpurun_out/q_bench.log
