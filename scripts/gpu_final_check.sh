# last check of the tree as committed: whole GPU suite, smoke, one bench line
python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-yardstick 2>gpurun_out/final_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'clk', d['clocks']['sm_mhz'], 'parity ok', d['parity']['ok'], 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))"
