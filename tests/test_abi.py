"""CPU-side checks of the boundary: the library builds, loads and exports every symbol declared in
include/gimb200.h; the packed blob matches the header's struct layout; the module mirrors the reference's
state_dict.  No compute is launched (there is no GPU here)."""
import ctypes
import os
import re
import struct

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from gim_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gimb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gimb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    from gim_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    # the test hooks live in their own library and header, not in the product ABI
    thdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gimb200_test.h")).read(), flags=re.S)
    tdecl = set(re.findall(r"\b(gimb_[a-z0-9_]+)\s*\(", thdr))
    assert tdecl == set(_lib.TEST_EXPORTS), tdecl ^ set(_lib.TEST_EXPORTS)
    tlib = _lib.load_test()
    for name in tdecl:
        assert hasattr(tlib, name) and not hasattr(lib, name), name


def test_abi_version_and_error_string(lib):
    assert lib.gimb_abi_version() == 1
    # a bad blob is refused with a message, not a crash (no GPU is touched before validation)
    h = ctypes.c_void_p()
    from gim_b200._lib import LoftrCfg
    cfg = LoftrCfg(0.2, 2, 0.1, 5)
    junk = ctypes.create_string_buffer(b"\0" * 64, 64)
    assert lib.gimb_loftr_create(junk, 64, ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"magic" in lib.gimb_last_error()


def test_blob_layout_matches_header():
    from gim_b200.weights import _ENTRY, _HEADER, load_gimw, DEFAULT_WEIGHTS, pack_loftr_blob, packed_tensors
    assert _HEADER.size == 32 and _ENTRY.size == 136  # sizeof(gimb_blob_header), sizeof(gimb_blob_entry)
    sd = load_gimw(DEFAULT_WEIGHTS)
    blob = pack_loftr_blob(sd)
    magic, ver, n, data_off, total = _HEADER.unpack_from(blob, 0)
    assert magic == 0x31304257424D4947 and ver == 1 and total == len(blob) and data_off % 256 == 0
    names = {}
    for i in range(n):
        name, ndim, s0, s1, s2, s3, _, off, nbytes = _ENTRY.unpack_from(blob, _HEADER.size + i * _ENTRY.size)
        names[name.rstrip(b"\0").decode()] = ((s0, s1, s2, s3)[:ndim], off, nbytes)
        assert off % 256 == 0
    pt = packed_tensors(sd)
    assert set(names) == set(pt)
    shape, off, nbytes = names["l2.0.c2.w"]
    assert shape == (128, 3, 3, 128)
    w = torch.frombuffer(bytearray(blob[data_off + off:data_off + off + nbytes]), dtype=torch.float32).view(shape)
    assert torch.equal(w, sd["backbone.encode.layer2.0.conv2.weight"].permute(0, 2, 3, 1))
    # folded BN == eval-mode BatchNorm
    x = torch.randn(4, 128, 5, 5)
    bn = torch.nn.functional.batch_norm(x, sd["backbone.encode.layer2.0.bn2.running_mean"],
                                        sd["backbone.encode.layer2.0.bn2.running_var"],
                                        sd["backbone.encode.layer2.0.bn2.weight"],
                                        sd["backbone.encode.layer2.0.bn2.bias"], False, 0.0, 1e-5)
    folded = x * pt["l2.0.c2.s"][None, :, None, None] + pt["l2.0.c2.b"][None, :, None, None]
    assert (bn - folded).abs().max() < 1e-5


def test_module_mirrors_reference_state_dict():
    from gim_b200 import LoFTR, get_default_config, load_default_weights
    m = LoFTR(get_default_config())
    sd = load_default_weights()
    ours = {k for k in m.state_dict() if not k.endswith("num_batches_tracked")}
    assert ours == set(sd)
    for k, v in m.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), k
    # the reference checkpoint keys carry a 'model.' prefix (loftr.py:93-99)
    res = m.load_state_dict({"model." + k: v for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys
    assert all(k.endswith("num_batches_tracked") for k in res.missing_keys)
    with pytest.raises(NotImplementedError):
        m.train()


def test_no_cpu_path():
    from gim_b200 import LoFTR, get_default_config
    m = LoFTR(get_default_config()).eval()
    c = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(dict(color0=c, color1=c, image0=c, image1=c))


def test_position_encoding_table_matches_oracle():
    from gim_b200.weights import position_encoding_table
    from oracle.loftr_oracle import position_encoding
    t = position_encoding_table(256, 6, 9)
    ref = position_encoding(256, 6, 9).permute(1, 2, 0).reshape(54, 256)
    assert torch.equal(t, ref)


# ----------------------------------------------------------------------------- gim_dkm boundary
def test_dkm_module_mirrors_reference_contract():
    """DKMv3(...) keeps the reference builder's signature, the caller-overwritten attributes (trainer/lightning.py:32-37)
    and - when /root/reference is present - exactly the reference's state_dict keys and shapes."""
    import os
    import sys
    from gim_b200 import DKMv3
    m = DKMv3(None, 672, 896, upsample_preds=True)
    assert (m.h_resized, m.w_resized, m.upsample_preds, m.symmetric, m.sample_mode) == (672, 896, True, True, "threshold_balanced")
    assert m.upsample_res == (1152, 1536) and m.sample_thresh == 0.05 and m.use_soft_mutual_nearest_neighbours is False
    m.h_resized, m.w_resized, m.upsample_res = 660, 880, (1152, 1536)   # what the ZEB harness does
    ref_root = os.environ.get("GIM_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(ref_root, "networks", "dkm", "models", "model_zoo", "DKMv3.py")):
        if ref_root not in sys.path:
            sys.path.insert(0, ref_root)
        from networks.dkm.models.model_zoo.DKMv3 import DKMv3 as RefDKMv3
        ref = RefDKMv3(None, 672, 896, upsample_preds=True).state_dict()
        ref = {k: v for k, v in ref.items() if "encoder.net.fc" not in k}   # callers strip fc (demo.py:358-363)
        ours = m.state_dict()
        assert set(ours) == set(ref)
        assert all(tuple(ours[k].shape) == tuple(ref[k].shape) for k in ref)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.match(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        DKMv3(None, 96, 128, symmetric=False).match(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 64, 64))


def test_dkm_blob_layout_and_folding():
    from gim_b200.dkm import packed_dkm_tensors
    from gim_b200.dkm_params import seeded_state_dict
    sd = seeded_state_dict(0)
    pt = packed_dkm_tensors(sd)
    assert pt["enc.l4.0.ds.w"].shape == (2048, 1, 1, 1024) and pt["proj.16.w"].shape == (512, 1, 1, 1024)
    assert pt["ref.16.b0.pw.w"].shape == (1377, 1, 1, 1377) and pt["ref.16.b0.dw_wt"].shape == (25, 1408)
    assert pt["ref.1.b0.dw_w"].shape == (24, 25) and pt["ref.1.out.w"].shape == (8, 1, 1, 24)       # 3 outputs padded to 8
    assert torch.equal(pt["ref.1.out.w"][3:], torch.zeros(5, 1, 1, 24)) and torch.equal(pt["ref.1.out.s"], torch.ones(8))
    # conv bias + BatchNorm folded into one affine: bn(conv(x) + bias) == s * conv(x) + b
    pre = "decoder.conv_refiner.8.hidden_blocks.2"
    x = torch.randn(2, 1137, 6, 7)
    y = torch.nn.functional.conv2d(x, sd[pre + ".0.weight"], sd[pre + ".0.bias"], padding=2, groups=1137)
    ref = torch.nn.functional.batch_norm(y, sd[pre + ".1.running_mean"], sd[pre + ".1.running_var"], sd[pre + ".1.weight"],
                                         sd[pre + ".1.bias"], False, 0.0, 1e-5)
    raw = torch.nn.functional.conv2d(x, sd[pre + ".0.weight"], None, padding=2, groups=1137)
    folded = raw * pt["ref.8.b3.dw_s"][None, :, None, None] + pt["ref.8.b3.dw_b"][None, :, None, None]
    assert (ref - folded).abs().max() < 1e-4
    wt = pt["ref.8.b3.dw_wt"]
    assert torch.equal(wt[:, :1137], sd[pre + ".0.weight"].reshape(1137, 25).t()) and torch.equal(wt[:, 1137:], torch.zeros(25, wt.shape[1] - 1137))
