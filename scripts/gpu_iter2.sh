# epilogue chunk-loop unroll A/B (experimental builds of the dev library), per-layer
SH="64,224,224,64,64,3 64,224,224,128,64,3 64,112,112,128,128,3 64,56,56,256,256,3 64,28,28,512,512,3 64,112,112,64,256,0 64,56,56,128,512,0 64,28,28,256,1024,0"
for lib in libugnet_dev.so libugnet_dev_u2.so libugnet_dev_u4.so; do
  echo "== $lib"
  UG_LIB_PATH=$PWD/unet-goolenet_b200/$lib UG_CONFIGS=v5 timeout 300 python scripts/conv_prof.py $SH 2>&1 | cut -c1-60,180-420 | tail -20
done
