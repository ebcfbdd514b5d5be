#!/usr/bin/env python
"""L2 -> SM bulk-tensor (TMA) load rate for the box shapes the GEMM kernel could use (see tma_probe in umma_gemm.cu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gim_b200 import _lib

NAMES = ["rows, 64-B box rows (BK=32, SW64)", "rows, 128-B box rows (BK=64, SW128)",
         "conv 8x16 patches, 64-B rows", "conv 8x16 patches, 128-B rows",
         "conv 4x32 patches, 64-B rows", "conv 4x32 patches, 128-B rows",
         "conv 2x64 patches, 64-B rows", "conv 2x64 patches, 128-B rows"]
lib = _lib.load_test()
for v, name in enumerate(NAMES):
    g = ctypes.c_float()
    _lib.check_test(lib.gimb_probe_tma(v, 4000, ctypes.byref(g), torch.cuda.current_stream().cuda_stream))
    print(f"{name:40s} {g.value:9.0f} GB/s  = {g.value * 1e9 / 148 / 1.9e9:6.1f} B/clk/SM @1.9 GHz", flush=True)
