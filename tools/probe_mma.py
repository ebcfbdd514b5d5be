#!/usr/bin/env python
"""Cycles per tcgen05.mma.kind::f16 (M = 128, K = 16) for the operand / accumulator patterns of the split-fp16 kernels
(see csrc/probe_mma.cu).  The floor is N / 2 cycles (128 x N x 16 MACs at 4096 MAC/clk/SM)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gim_b200 import _lib

lib = _lib.load_test()
lib.gimb_probe_mma.restype = ctypes.c_int
lib.gimb_probe_mma.argtypes = [ctypes.c_int] * 5 + [ctypes.POINTER(ctypes.c_float)] * 3 + [ctypes.c_void_p]
NAMES = {0: "same operands, 1 accumulator", 1: "walk k, hi*hi only, 1 accumulator", 2: "split scheme, X/X/Y accumulators",
         3: "split scheme, 1 accumulator", 4: "split scheme, 3 accumulators"}
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
for grid, ns, vs, lws in ((1, (64, 128, 256), range(5), (0,)), (148, (128, 256), (2, 3), (0, 4, 8))):
    for n in ns:
        for v in vs:
            if (v == 4 and n > 128) or (v == 2 and n > 128):
                continue
            for lw in lws:
                c, ms, lb = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
                _lib.check_test(lib.gimb_probe_mma(v, n, iters, grid, lw, ctypes.byref(c), ctypes.byref(ms), ctypes.byref(lb),
                                                   torch.cuda.current_stream().cuda_stream))
                mmas = iters * 12
                tf = grid * mmas * 128 * n * 16 * 2 / (ms.value * 1e-3) / 1e12
                print(f"grid {grid:3d} N {n:3d} {NAMES[v]:38s} + {lw} TMEM-reading warps ({lb.value:6.1f} B/clk/SM read): {c.value:7.1f} cyc/MMA "
                      f"(floor {n // 2:3d})  {tf:7.1f} TFLOP/s executed", flush=True)
