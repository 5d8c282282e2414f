#!/usr/bin/env python3
"""Phase timing (SM clock cycles) of one 64x64 diagonal-block factorization (potrf_block.cuh)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import numpy as np
from mrcal_b200 import _capi
out = np.zeros(64, np.int64)
assert _capi.lib.mrcal_b200_debug_potrf_stamps(out.ctypes.data_as(C.c_void_p))
t0 = out[30]
rel = lambda i: int(out[i] - t0) if out[i] else None
print("kernel start..end:", rel(31), "cycles; block start", rel(0), "block end", rel(20))
for p in range(4):
    print(f"panel {p}: warp0 start {rel(1+4*p)} factor_done {rel(2+4*p)} published {rel(3+4*p)} | "
          f"bulk: got {rel(33+6*p)} B1done {rel(34+6*p)} B2start {rel(35+6*p)} diag_tile_done {rel(36+6*p)} B2done {rel(37+6*p)}")
print("warp1 start", rel(32), "end", rel(60))
