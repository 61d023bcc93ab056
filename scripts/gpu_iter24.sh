# plan-time autotune with the pair forms among the candidates: parity and throughput of the tuned program
UG_AUTOTUNE=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-yardstick --parity-images 32 2>gpurun_out/i24_err.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('UG_AUTOTUNE=1', round(d['value']), 'img/s', round(d['ms_per_step'],2), 'ms/step, e2e', round(d['e2e']['value']), 'parity ok', d['parity']['ok'], 'tuned ops', d['config'].get('autotuned_conv_ops'))"
tail -2 gpurun_out/i24_err.log
