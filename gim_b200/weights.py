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


# ----------------------------------------------------------------------------- packed blob for the C ABI
BLOB_MAGIC = 0x31304257424D4947  # "GIMBWB01"
_ENTRY = struct.Struct("<96sI4IIQQ")  # name, ndim, shape[4], reserved, offset, nbytes  (include/gimb200.h)
_HEADER = struct.Struct("<QIIQQ")     # magic, version, n_entries, data_offset, total_bytes


def _fold_bn(sd, pre, eps=1e-5):
    """Eval-mode BatchNorm as y = x * s + b, folded in fp32 the way ATen's CPU kernel does
    (alpha = weight / sqrt(var + eps), beta = bias - mean * alpha)."""
    inv_std = 1.0 / torch.sqrt(sd[pre + ".running_var"].float() + eps)
    s = sd[pre + ".weight"].float() * inv_std
    b = sd[pre + ".bias"].float() - sd[pre + ".running_mean"].float() * s
    return s, b


def _khwc(w):
    """conv weight OIHW -> [O, kh, kw, I] (the K order of the NHWC implicit GEMM)."""
    return w.float().permute(0, 2, 3, 1).contiguous()


def packed_tensors(state_dict):
    """state_dict (reference key names) -> ordered {packed name: fp32 tensor} consumed by libgimb200."""
    sd = strip_prefix(state_dict)
    out = {}

    def conv(name, wkey, bnkey=None):
        out[name + ".w"] = _khwc(sd[wkey])
        if bnkey is not None:
            out[name + ".s"], out[name + ".b"] = _fold_bn(sd, bnkey)

    enc = "backbone.encode"
    conv("stem", enc + ".conv1.weight", enc + ".bn1")
    for li, nblk in ((1, 3), (2, 4), (3, 6)):
        for bi in range(nblk):
            pre, name = f"{enc}.layer{li}.{bi}", f"l{li}.{bi}"
            for ci in (1, 2, 3):
                conv(f"{name}.c{ci}", f"{pre}.conv{ci}.weight", f"{pre}.bn{ci}")
            if f"{pre}.downsample.0.weight" in sd:
                conv(f"{name}.ds", f"{pre}.downsample.0.weight", f"{pre}.downsample.1")
    conv("fpn.l3out", "backbone.layer3_outconv.weight")
    conv("fpn.l2out", "backbone.layer2_outconv.weight")
    conv("fpn.l2c1", "backbone.layer2_outconv2.0.weight", "backbone.layer2_outconv2.1")
    conv("fpn.l2c2", "backbone.layer2_outconv2.3.weight")
    conv("fpn.l1out", "backbone.layer1_outconv.weight")
    conv("fpn.l1c1", "backbone.layer1_outconv2.0.weight", "backbone.layer1_outconv2.1")
    conv("fpn.l1c2", "backbone.layer1_outconv2.3.weight")
    for mod, short, nl in (("loftr_coarse", "coarse", 8), ("loftr_fine", "fine", 2)):
        for i in range(nl):
            pre, name = f"{mod}.layers.{i}", f"{short}.{i}"
            out[name + ".q"] = sd[pre + ".q_proj.weight"].float().contiguous()
            out[name + ".kv"] = torch.cat([sd[pre + ".k_proj.weight"], sd[pre + ".v_proj.weight"]], 0).float().contiguous()
            out[name + ".merge"] = sd[pre + ".merge.weight"].float().contiguous()
            out[name + ".mlp0"] = sd[pre + ".mlp.0.weight"].float().contiguous()
            out[name + ".mlp2"] = sd[pre + ".mlp.2.weight"].float().contiguous()
            out[name + ".n1g"], out[name + ".n1b"] = sd[pre + ".norm1.weight"].float(), sd[pre + ".norm1.bias"].float()
            out[name + ".n2g"], out[name + ".n2b"] = sd[pre + ".norm2.weight"].float(), sd[pre + ".norm2.bias"].float()
    return out


def pack_loftr_blob(state_dict):
    """-> bytes: the blob `gimb_loftr_create` takes (layout: include/gimb200.h)."""
    tensors = packed_tensors(state_dict)
    entries, chunks, off = [], [], 0
    for name, t in tensors.items():
        a = t.detach().cpu().contiguous().numpy().astype(np.float32, copy=False)
        shape = list(a.shape) + [0] * (4 - a.ndim)
        entries.append(_ENTRY.pack(name.encode(), a.ndim, *shape, 0, off, a.nbytes))
        raw = a.tobytes()
        pad = (-len(raw)) % 256
        chunks.append(raw + b"\0" * pad)
        off += len(raw) + pad
    table = b"".join(entries)
    data_offset = (_HEADER.size + len(table) + 255) // 256 * 256
    head = _HEADER.pack(BLOB_MAGIC, 1, len(entries), data_offset, data_offset + off)
    return head + table + b"\0" * (data_offset - _HEADER.size - len(table)) + b"".join(chunks)


def position_encoding_table(d_model, hc, wc):
    """The reference's PositionEncodingSine buffer with temp_bug_fix=False (loftr.py:22-24,
    position_encoding.py:22-37) for an hc x wc coarse map, token-major [hc*wc, d_model].
    `-math.log(10000.0) / d_model//2` evaluates to -1.0; positions are 1-based (cumsum of ones)."""
    import math
    y = torch.ones(hc, wc).cumsum(0).float().unsqueeze(0)
    x = torch.ones(hc, wc).cumsum(1).float().unsqueeze(0)
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))[:, None, None]
    pe = torch.zeros(d_model, hc, wc)
    pe[0::4] = torch.sin(x * div)
    pe[1::4] = torch.cos(x * div)
    pe[2::4] = torch.sin(y * div)
    pe[3::4] = torch.cos(y * div)
    return pe.permute(1, 2, 0).reshape(hc * wc, d_model).contiguous()
