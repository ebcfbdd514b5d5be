"""Helpers shared by the parity tests: load golden cases, compare outputs."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: coarse ids bit-exact; sub-pixel coordinates and confidences within 1e-3 abs.
TOL_KPTS = 1e-3
TOL_CONF = 1e-3

CASES = ["tiny_64x96", "noise_96x128", "nomatch_flat_96x128", "diffsize_256x320_320x256",
         "masked_scaled_b2_256x320", "small_b2_240x320", "demo_a_480x640", "synth_b2_480x640"]


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    data = {
        "color0": torch.from_numpy(z["color0_u8"]).float() / 255.0,
        "color1": torch.from_numpy(z["color1_u8"]).float() / 255.0,
    }
    data["image0"], data["image1"] = data["color0"], data["color1"]
    for k in ("mask0", "mask1"):
        if k in z.files:
            data[k] = torch.from_numpy(z[k])
    for k in ("scale0", "scale1"):
        if k in z.files:
            data[k] = torch.from_numpy(z[k])
    gold = {k: torch.from_numpy(z[k]) for k in z.files
            if k not in ("color0_u8", "color1_u8", "mask0", "mask1", "scale0", "scale1")}
    return data, gold


def assert_matches_equal(out, gold, what=""):
    """ids bit-exact (and in torch.where order), floats within the north_star tolerance."""
    for k in ("b_ids", "i_ids", "j_ids"):
        a, b = out[k].cpu().long(), gold[k].long()
        assert a.shape == b.shape, f"{what}{k}: M differs: {tuple(a.shape)} vs {tuple(b.shape)}"
        assert torch.equal(a, b), f"{what}{k}: ids differ at {(a != b).nonzero()[:5].flatten().tolist()}"
    assert torch.equal(out["m_bids"].cpu().long(), gold["m_bids"].long())
    if gold["b_ids"].numel() == 0:
        for k in ("mkpts0_f", "mkpts1_f", "mconf"):
            assert out[k].shape[0] == 0
        return {}
    errs = {}
    for k, tol in (("mconf", TOL_CONF), ("mkpts0_c", 0.0), ("mkpts1_c", 0.0), ("mkpts0_f", 0.0),
                   ("mkpts1_f", TOL_KPTS), ("expec_f", TOL_KPTS)):
        a, b = out[k].cpu().float(), gold[k].float()
        assert a.shape == b.shape, f"{what}{k}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
        e = (a - b).abs().max().item()
        errs[k] = e
        assert e <= tol, f"{what}{k}: max abs err {e} > {tol}"
    return errs
