"""Probe (next-round design): does cuTensorMapEncodeTiled accept a tensor map whose rows OVERLAP
(pixel stride 32 B under a 64-element = 128-byte inner dimension)?  Prints the CUresult of a normal and an overlapping
encode on a device buffer."""
import torch
from cuda.bindings import driver as cu

torch.cuda.init()
buf = torch.zeros((4, 115, 115, 16), device="cuda", dtype=torch.bfloat16)
T = cu.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
args = (cu.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, cu.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B,
        cu.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_128B, cu.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
u64, u32 = cu.cuuint64_t, cu.cuuint32_t
for name, dims, strides, box in (
        ("plain   [16ch, 115, 115, 4]", (16, 115, 115, 4), (32, 115 * 32, 115 * 115 * 32), (16, 16, 8, 1)),
        ("overlap [64 (4px x 16ch), 112, 112, 4], pixel stride 32 B", (64, 112, 112, 4), (32, 115 * 32, 115 * 115 * 32),
         (64, 16, 8, 1))):
    r = cu.cuTensorMapEncodeTiled(T, 4, buf.data_ptr(), [u64(v) for v in dims], [u64(v) for v in strides],
                                  [u32(v) for v in box], [u32(1)] * 4, *args)
    print(name, "->", r[0])
