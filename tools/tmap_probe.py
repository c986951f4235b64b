"""Probe: does cuTensorMapEncodeTiled accept overlapping (sliding-window) strides?"""
import torch
from cuda.bindings import driver as drv
x = torch.zeros(2 * 4 * 16 * 392 * 8, dtype=torch.bfloat16, device="cuda")
u64, u32 = drv.cuuint64_t, drv.cuuint32_t
def enc(dims, strides, box, es=(1, 1, 1, 1, 1)):
    r = drv.cuTensorMapEncodeTiled(drv.CUtensorMapDataType.CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, x.data_ptr(),
                                   [u64(d) for d in dims], [u64(s) for s in strides], [u32(b) for b in box], [u32(e) for e in es],
                                   drv.CUtensorMapInterleave.CU_TENSOR_MAP_INTERLEAVE_NONE, drv.CUtensorMapSwizzle.CU_TENSOR_MAP_SWIZZLE_128B,
                                   drv.CUtensorMapL2promotion.CU_TENSOR_MAP_L2_PROMOTION_L2_256B, drv.CUtensorMapFloatOOBfill.CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
    return r[0]
pitch = 392 * 16
print("dense      ", enc((8, 392, 16, 4, 2), (16, pitch, 16 * pitch, 4 * 16 * pitch), (8, 64, 1, 1, 1)))
print("overlap    ", enc((64, 192, 16, 4, 2), (32, pitch, 16 * pitch, 4 * 16 * pitch), (64, 64, 1, 2, 1)))
print("overlap+es ", enc((64, 192, 16, 4, 2), (32, pitch, 16 * pitch, 4 * 16 * pitch), (64, 64, 4, 1, 1), (1, 1, 2, 1, 1)))
