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
