#!/usr/bin/env python
"""Single-pair latency of the gim_loftr forward (BASELINE config 1 = the demo pair after demo.py's pre-processing, 1000x1000;
plus 480x640), device-resident inputs and end to end through the uint8 host entry.  Median of 20 calls after 5 warm-ups."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gim_b200 import LoFTR, get_default_config, load_default_weights, synth
from tests.goldens import load_case

m = LoFTR(get_default_config()); m.load_state_dict(load_default_weights()); m = m.eval().cuda()
out = {}
for name in ("demo_a_1000x1000", "synth_480x640"):
    if name.startswith("demo"):
        data, _ = load_case(name)
        c0, c1 = data["color0"], data["color1"]
    else:
        c0, c1 = synth.make_pairs(1, 480, 640, first=0)
    d0, d1 = c0.cuda(), c1.cuda()
    u0 = torch.round(c0 * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    u1 = torch.round(c1 * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    def dev():
        d = dict(color0=d0, color1=d1, image0=d0, image1=d1); m(d); return d
    def host():
        d = dict(color0_u8=u0, color1_u8=u1); m.forward_u8(d); return d
    res = {}
    for label, fn in (("device_ms", dev), ("e2e_u8_ms", host)):
        for _ in range(5): fn()
        ts = []
        for _ in range(20):
            torch.cuda.synchronize(); t = time.perf_counter(); d = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
        ts.sort(); res[label] = ts[len(ts) // 2]; res["matches"] = int(d["b_ids"].numel())
    out[name] = res
print(json.dumps({"metric": "single-pair latency, gim_loftr", "unit": "ms (median of 20)", "results": out}))
