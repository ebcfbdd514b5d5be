"""Pins the oracle (oracle/loftr_oracle.py) to golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py) and - where /root/reference exists - to the live reference module."""
import os

import pytest
import torch

from gim_b200.weights import DEFAULT_WEIGHTS, load_gimw
from oracle import loftr_oracle
from oracle.ref_import import reference_available
from tests.goldens import CASES, assert_matches_equal, load_case

FAST = [c for c in CASES if "480x640" not in c]


@pytest.fixture(scope="module")
def weights():
    return load_gimw(DEFAULT_WEIGHTS)


@pytest.mark.parametrize("case", FAST)
def test_oracle_matches_reference_golden(weights, case):
    data, gold = load_case(case)
    out = loftr_oracle.loftr_forward(weights, data, return_intermediates=True)
    errs = assert_matches_equal(out, gold, what=case + ": ")
    inter = out["_inter"]
    if "inter_feat_c_backbone" in gold:
        n = data["color0"].shape[0]
        ref_c = gold["inter_feat_c_backbone"]
        got = torch.cat([inter["feat_c0_backbone"], inter["feat_c1_backbone"]], 0)
        assert (got - ref_c).abs().max() < 2e-5
        got_f = torch.cat([inter["feat_f0"], inter["feat_f1"]], 0)
        assert (got_f - gold["inter_feat_f"]).abs().max() < 2e-5
    if "inter_feat_c0" in gold:
        assert (inter["feat_c0"] - gold["inter_feat_c0"]).abs().max() < 5e-5
        assert (inter["feat_c1"] - gold["inter_feat_c1"]).abs().max() < 5e-5
    if "inter_fine_win0" in gold and gold["b_ids"].numel():
        assert (inter["fine_win0"] - gold["inter_fine_win0"]).abs().max() < 5e-5
        assert (inter["fine_win1"] - gold["inter_fine_win1"]).abs().max() < 5e-5
    # the restatement is the same arithmetic, so it is far tighter than the north_star tolerance
    for k, e in errs.items():
        assert e < 2e-4, (k, e)


@pytest.mark.slow
def test_oracle_matches_reference_golden_480x640(weights):
    data, gold = load_case("demo_a_480x640")
    out = loftr_oracle.loftr_forward(weights, data)
    assert_matches_equal(out, gold, what="demo_a_480x640: ")


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference(weights):
    from oracle.ref_import import load_reference_loftr
    from gim_b200 import synth
    model = load_reference_loftr()
    c0, c1 = synth.make_pairs(1, 128, 160, first=5)
    data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
    with torch.no_grad():
        model(data)
    out = loftr_oracle.loftr_forward(weights, dict(color0=c0, color1=c1))
    assert_matches_equal(out, {k: data[k] for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mconf", "mkpts0_c",
                                                     "mkpts1_c", "mkpts0_f", "mkpts1_f", "expec_f")})
    assert (out["mconf"] - data["mconf"]).abs().max() < 1e-5


def test_position_encoding_buggy_branch():
    """temp_bug_fix=False: div_term = exp(-k), k = 0, 2, ... (position_encoding.py:29)."""
    pe = loftr_oracle.position_encoding(256, 4, 6)
    assert pe.shape == (256, 4, 6)
    assert torch.allclose(pe[0, 0, :], torch.sin(torch.arange(1, 7).float()))
    assert torch.allclose(pe[4, 0, :], torch.sin(torch.arange(1, 7).float() * torch.exp(torch.tensor(-2.0))))
    assert torch.allclose(pe[3, :, 0], torch.cos(torch.arange(1, 5).float()))
