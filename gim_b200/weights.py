"""Weight containers for the B200 gim_loftr path.

Two formats live here (both host-side, product code):

1. `.gimw` - an on-disk container of the *raw* checkpoint tensors (checkpoint key names minus the
   `model.` / `matcher.` prefix that `networks/loftr/loftr.py:93-99` strips).  It can be sharded
   (`name.gimw.0`, `name.gimw.1`, ...) so no file exceeds 50 MB.
2. the *packed blob* handed to `gimb_loftr_create()` (layout in include/gimb200.h): BN folded to
   per-channel scale/bias, conv weights re-ordered OIHW -> O,kh,kw,I (NHWC implicit-GEMM K order),
   k/v projection weights fused, PE tables omitted (the shim passes them per call shape).
"""
import glob
import json
import os
import struct

import numpy as np
import torch

GIMW_MAGIC = b"GIMW0001"


# ----------------------------------------------------------------------------- .gimw container
def strip_prefix(state_dict):
    """Same key normalisation as the reference's LoFTR.load_state_dict (loftr.py:93-99)."""
    out = {}
    for k, v in state_dict.items():
        for p in ("model.", "matcher."):
            if k.startswith(p):
                k = k[len(p):]
                break
        out[k] = v
    return out


def save_gimw(state_dict, path, shard_bytes=48 << 20):
    sd = {k: v for k, v in strip_prefix(state_dict).items() if not k.endswith("num_batches_tracked")}
    index, off = [], 0
    for k, v in sd.items():
        a = v.detach().cpu().contiguous().to(torch.float32).numpy()
        index.append({"name": k, "shape": list(a.shape), "offset": off})
        off += a.nbytes
    head = json.dumps({"tensors": index, "total": off}).encode()
    payload = GIMW_MAGIC + struct.pack("<Q", len(head)) + head
    payload += b"".join(v.detach().cpu().contiguous().to(torch.float32).numpy().tobytes() for v in sd.values())
    for old in glob.glob(path + ".*"):
        os.remove(old)
    nshard = (len(payload) + shard_bytes - 1) // shard_bytes
    for s in range(nshard):
        with open(f"{path}.{s}", "wb") as f:
            f.write(payload[s * shard_bytes:(s + 1) * shard_bytes])
    return nshard


def load_gimw(path):
    """-> {name: torch.float32 tensor}.  `path` is the un-suffixed name; shards are concatenated."""
    shards = sorted(glob.glob(path + ".*"), key=lambda p: int(p.rsplit(".", 1)[1]))
    if not shards:
        raise FileNotFoundError(f"no weight shards at {path}.N")
    buf = b"".join(open(p, "rb").read() for p in shards)
    if buf[:8] != GIMW_MAGIC:
        raise ValueError("not a .gimw container")
    (hl,) = struct.unpack("<Q", buf[8:16])
    meta = json.loads(buf[16:16 + hl])
    base = 16 + hl
    out = {}
    for t in meta["tensors"]:
        n = int(np.prod(t["shape"])) if t["shape"] else 1
        a = np.frombuffer(buf, dtype=np.float32, count=n, offset=base + t["offset"]).reshape(t["shape"])
        out[t["name"]] = torch.from_numpy(a.copy())
    return out


DEFAULT_WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                               "weights", "gim_loftr_50h.gimw")
