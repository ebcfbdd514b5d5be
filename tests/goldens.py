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


# Cases whose inputs are regenerated from a recipe (gim_b200.synth is deterministic) instead of being stored:
#   kitti_pad_b2_1240   ZEB KITTI geometry (TEST_GIM_LOFTR.sh:4, datasets/utils.py:80-126): 376x1240 content zero-padded
#                       to 1240x1240 with mask0/mask1 and scale0/scale1, batch 2 (L = S = 155*155 = 24025)
#   eth3d_b1_1064x1600  ETH3D geometry (TEST_GIM_LOFTR.sh:10-11): 1600-wide, batch 1 (L = S = 26600)
#   synth_p13_p31       pairs 13 and 31 of the headline batch (bench.py: synth.make_pairs(32, 480, 640, first=0))
RECIPES = ["kitti_pad_b2_1240", "eth3d_b1_1064x1600", "synth_p13_p31"]
BIG_CASES = ["demo_a_1000x1000", "kitti_pad_b2_1240", "eth3d_b1_1064x1600"]


def build_recipe(name):
    from gim_b200 import synth
    if name == "kitti_pad_b2_1240":
        a0, a1 = synth.make_pairs(2, 376, 1240, first=6)
        c0 = torch.zeros(2, 3, 1240, 1240)
        c1 = torch.zeros(2, 3, 1240, 1240)
        c0[:, :, :376] = a0
        c1[:, :, :376] = a1
        m0 = torch.zeros(2, 155, 155, dtype=torch.bool)
        m0[:, :47] = True
        data = dict(color0=c0, color1=c1, mask0=m0, mask1=m0.clone(),
                    scale0=torch.tensor([[1242 / 1240, 375 / 376]] * 2), scale1=torch.tensor([[1241 / 1240, 376 / 376]] * 2))
    elif name == "eth3d_b1_1064x1600":
        c0, c1 = synth.make_pairs(1, 1064, 1600, first=9)
        data = dict(color0=c0, color1=c1)
    elif name == "synth_p13_p31":
        a, b = synth.make_pairs(1, 480, 640, first=13), synth.make_pairs(1, 480, 640, first=31)
        data = dict(color0=torch.cat([a[0], b[0]]), color1=torch.cat([a[1], b[1]]))
    else:
        raise KeyError(name)
    data["image0"], data["image1"] = data["color0"], data["color1"]
    return data


def load_case(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    if name in RECIPES:
        return build_recipe(name), {k: torch.from_numpy(z[k]) for k in z.files}
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
