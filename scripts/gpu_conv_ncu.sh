# ncu full captures (source view) of single conv launches: args = list of "shape variant [mode]" strings
i=0
for cfg in "$@"; do
  i=$((i+1))
  python scripts/conv_one.py $cfg > gpurun_out/one_$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv -s 2 -c 1 -o gpurun_out/cap_$i python scripts/conv_one.py $cfg > gpurun_out/ncu_one_$i.log 2>&1
  echo "ncu $cfg rc=$?"
  ncu -i gpurun_out/cap_$i.ncu-rep --page raw --csv > gpurun_out/cap_$i.raw.csv 2>/dev/null
  ncu -i gpurun_out/cap_$i.ncu-rep --page source --csv > gpurun_out/cap_$i.source.csv 2>/dev/null
  rm -f gpurun_out/cap_$i.ncu-rep
done
