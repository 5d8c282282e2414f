#!/usr/bin/env python3
"""Warm device timings of the dense Cholesky and of its kernel kinds. usage: time_cholesky.py n [n...]"""
import ctypes as C
import sys
sys.path.insert(0, ".")
from mrcal_b200 import _capi
f = _capi.lib.mrcal_b200_debug_time_cholesky
f.restype = C.c_double
for n in [int(a) for a in sys.argv[1:]] or [1268, 4820]:
    full_g = f(n, 10, 15, 1)
    full_d = f(n, 10, 15, 0)
    parts = {name: f(n, 10, k, 1) for name, k in (("potrf_diag", 1), ("trsm", 2), ("syrk_panel", 4), ("syrk_trailing", 8))}
    gf = n ** 3 / 3 / 1e9
    print(f"n={n}: multi-kernel graph {full_g:.3f} ms ({gf / full_g:.1f} TFLOP/s, GFLOP={gf:.2f}), direct launches {full_d:.3f} ms; "
          + ", ".join(f"{k} {v:.3f}" for k, v in parts.items()))
    if n <= 8192:
        fill = f(n, 10, 33, 0)
        df = f(n, 10, 32, 0) - fill
        print(f"n={n}: persistent factorization {df:.3f} ms ({gf / df:.1f} TFLOP/s); backward substitution: persistent {f(n, 20, 64, 0):.4f} ms, "
              f"multi-kernel graph {f(n, 20, 128, 0):.4f} ms")
