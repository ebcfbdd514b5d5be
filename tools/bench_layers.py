#!/usr/bin/env python
"""Per-layer timing of the tcgen05 GEMM at the N = 32 pairs (64 images) @ 480x640 workload shapes.
Prints achieved MMA TFLOP/s (3 MMAs per logical MAC), algorithmic TFLOP/s and effective HBM GB/s per layer."""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gim_b200 import _lib

B = int(os.environ.get("IMGS", "64"))
LAYERS = [
    # name, B, H, W, C1, C2, Cout, k, stride, flags, act
    ("l1.0.c1 1x1 64->64", B, 240, 320, 64, 0, 64, 1, 1, 1 | 8, 1),
    ("l1.x.c2 3x3 64->64", B, 240, 320, 64, 0, 64, 3, 1, 1 | 8, 1),
    ("l1.0.ds 1x1 64->256 f32", B, 240, 320, 64, 0, 256, 1, 1, 1 | 4, 0),
    ("l1.x.c3 1x1 64->256 +res", B, 240, 320, 64, 0, 256, 1, 1, 1 | 2 | 4 | 8, 1),
    ("l1.x.c3p 1x1 64->256 +res planes", B, 240, 320, 64, 0, 256, 1, 1, 1 | 8 | 16, 1),
    ("l1.x.c3n 1x1 64->256 planes, no res", B, 240, 320, 64, 0, 256, 1, 1, 1 | 8, 1),
    ("l2.x.c3p 1x1 128->512 +res planes", B, 120, 160, 128, 0, 512, 1, 1, 1 | 8 | 16, 1),
    ("l3.x.c3p 1x1 256->1024 +res planes", B, 60, 80, 256, 0, 1024, 1, 1, 1 | 8 | 16, 1),
    ("l1.x.c1 1x1 256->64", B, 240, 320, 256, 0, 64, 1, 1, 1 | 8, 1),
    ("l2.0.c2 3x3s2 128->128", B, 240, 320, 128, 0, 128, 3, 2, 1 | 8, 1),
    ("l2.x.c2 3x3 128->128", B, 120, 160, 128, 0, 128, 3, 1, 1 | 8, 1),
    ("l2.x.c3 1x1 128->512 +res", B, 120, 160, 128, 0, 512, 1, 1, 1 | 2 | 4 | 8, 1),
    ("l2.x.c1 1x1 512->128", B, 120, 160, 512, 0, 128, 1, 1, 1 | 8, 1),
    ("l3.x.c2 3x3 256->256", B, 60, 80, 256, 0, 256, 3, 1, 1 | 8, 1),
    ("l3.x.c3 1x1 256->1024 +res", B, 60, 80, 256, 0, 1024, 1, 1, 1 | 2 | 4 | 8, 1),
    ("l3.x.c1 1x1 1024->256", B, 60, 80, 1024, 0, 256, 1, 1, 1 | 8, 1),
    ("fpn.l2c1 3x3 256->256", B, 120, 160, 256, 0, 256, 3, 1, 1 | 8, 2),
    ("fpn.l1out 1x1 256->196 f32", B, 240, 320, 256, 0, 196, 1, 1, 4, 0),
    ("fpn.l1c1 3x3 196->196", B, 240, 320, 196, 0, 196, 3, 1, 1 | 8, 2),
    ("fpn.l1c2 3x3 196->128 f32", B, 240, 320, 196, 0, 128, 3, 1, 4, 0),
    ("coarse q 256->256 f32", 1, B * 4800, 1, 256, 0, 256, 1, 1, 4, 3),
    ("coarse mlp0 512->512", 1, B * 4800, 1, 256, 256, 512, 1, 1, 8, 1),
    ("coarse mlp2 512->256 f32", 1, B * 4800, 1, 512, 0, 256, 1, 1, 4, 0),
    ("fine q 128->128 f32 (M=112932)", 1, 2 * 112932 * 25 // 8 * 8, 1, 128, 0, 128, 1, 1, 4, 3),
]

def main():
    lib = _lib.load_test()
    only = sys.argv[1:]
    for (name, b, h, w, c1, c2, co, k, s, flags, act) in LAYERS:
        if only and not any(o in name for o in only):
            continue
        ms = ctypes.c_float()
        _lib.check_test(lib.gimb_bench_layer(b, h, w, c1, c2, co, k, s, flags, act, 5, ctypes.byref(ms),
                                        torch.cuda.current_stream().cuda_stream))
        pad = k // 2
        oh, ow = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        M = b * oh * ow
        K = k * k * (c1 + c2)
        gflop = 2.0 * M * co * K / 1e9
        byt = b * h * w * (c1 + c2) * 4 + M * co * ((4 if flags & 2 else 0) + (4 if flags & 4 else 0) + (4 if flags & 8 else 0) + (4 if flags & 16 else 0))
        t = ms.value / 1e3
        print(f"{name:34s} {ms.value:8.3f} ms  alg {gflop / t / 1e3:7.1f} TF/s  mma {3 * gflop / t / 1e3:7.1f} TF/s  "
              f"hbm {byt / t / 1e9:7.0f} GB/s", flush=True)

if __name__ == "__main__":
    main()
