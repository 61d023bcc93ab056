"""ugnet_b200 — B200-native engine for the UNet -> bbox crop -> GoogLeNet hot path of BY-Elysia/UNet-GooLeNet.

Import as `import ugnet_b200` (see ugnet_b200.py at the repo root).  Sub-modules:
  engine   ctypes binding of libugnet.so (the C ABI in include/ugnet.h)
  pack     BN folding and bf16 K-major weight packing
  nets     drop-in shells for the reference's `nets` package (UNetTaskAligWeight, ...)
  googlenet, util.roi, pipeline, infer, dist — the rest of the host-side mirror of the reference path
"""
__version__ = "0.1.0"
