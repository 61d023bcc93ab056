# per-role counters of the multi-issuer kernel + ncu full captures of single conv launches (small reports)
timeout 600 python scripts/conv_prof.py 64,224,224,64,64,3 64,224,224,128,64,3 64,112,112,128,128,3 64,56,56,256,256,3 > gpurun_out/conv_sweep2.log 2>&1; cut -c1-700 gpurun_out/conv_sweep2.log
i=0
for cfg in "64,224,224,64,64,3 5" "64,224,224,64,64,3 3" "64,112,112,128,128,3 5" "64,112,112,128,128,3 3"; do
  i=$((i+1))
  python scripts/conv_one.py $cfg > gpurun_out/one_$i.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 2 -c 1 -o gpurun_out/r01c_conv_$i python scripts/conv_one.py $cfg > gpurun_out/ncu_one_$i.log 2>&1
  echo "ncu $cfg rc=$?"
  ncu -i gpurun_out/r01c_conv_$i.ncu-rep --page raw --csv > gpurun_out/r01c_conv_$i.raw.csv 2>/dev/null
done
ls -la gpurun_out
