timeout 120 python scripts/probe_overlap_conv.py 2>&1 | tail -4
UG_CONFIGS=convt timeout 300 python scripts/conv_prof.py 128,14,14,512,2048,0 128,28,28,256,1024,0 128,56,56,128,512,0 128,112,112,64,256,0 2>&1 | cut -c1-100 | tail -16
