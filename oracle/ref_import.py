"""TEST INFRASTRUCTURE: import the unmodified reference LoFTR from /root/reference.

The reference (`networks/loftr/loftr.py`) needs torch + einops + two kornia symbols
(`networks/loftr/utils/fine_matching.py:5-6`).  kornia is not installed here, so the two
symbols are provided as in-memory stub modules that restate kornia 0.6.10's published
semantics (`environment.yaml:21`):

* `create_meshgrid(h, w, normalized_coordinates=True, device)` -> [1, h, w, 2], last dim (x, y),
  x = linspace(-1, 1, w), y = linspace(-1, 1, h);
* `dsnt.spatial_expectation2d(p[B, C, h, w], normalized_coordinates=True)` -> [B, C, 2] =
  (sum x*p, sum y*p).

Only usable where /root/reference exists (the build container); never on the GPU box.
"""
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("GIM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks", "loftr", "loftr.py"))


def _install_kornia_stub():
    if "kornia" in sys.modules and not getattr(sys.modules["kornia"], "_gimb_stub", False):
        return  # a real kornia is importable; use it

    def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=torch.float32):
        xs = torch.linspace(0, width - 1, width, device=device, dtype=dtype)
        ys = torch.linspace(0, height - 1, height, device=device, dtype=dtype)
        if normalized_coordinates:
            xs = (xs / (width - 1) - 0.5) * 2
            ys = (ys / (height - 1) - 0.5) * 2
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        return torch.stack([gx, gy], dim=-1)[None]

    def spatial_expectation2d(inp, normalized_coordinates=True):
        b, c, h, w = inp.shape
        grid = create_meshgrid(h, w, normalized_coordinates, inp.device).to(inp.dtype)
        pos_x = grid[..., 0].reshape(-1)
        pos_y = grid[..., 1].reshape(-1)
        flat = inp.reshape(b, c, -1)
        ex = torch.sum(pos_x * flat, -1, keepdim=True)
        ey = torch.sum(pos_y * flat, -1, keepdim=True)
        return torch.cat([ex, ey], -1).reshape(b, c, 2)

    def mod(name):
        m = types.ModuleType(name)
        m._gimb_stub = True
        sys.modules[name] = m
        return m

    kornia = mod("kornia")
    geometry = mod("kornia.geometry")
    subpix = mod("kornia.geometry.subpix")
    dsnt = mod("kornia.geometry.subpix.dsnt")
    utils = mod("kornia.utils")
    grid = mod("kornia.utils.grid")
    kornia.geometry, kornia.utils = geometry, utils
    geometry.subpix = subpix
    subpix.dsnt = dsnt
    utils.grid = grid
    dsnt.spatial_expectation2d = spatial_expectation2d
    grid.create_meshgrid = create_meshgrid
    utils.create_meshgrid = create_meshgrid


def default_loftr_config():
    """`lower_config(get_cfg_defaults())['loftr']` written literally (yacs is absent here);
    values from networks/loftr/config.py:7-46."""
    return {
        "weight": None,
        "backbone_type": "ResNetFPN",
        "resolution": (8, 2),
        "fine_window_size": 5,
        "fine_concat_coarse_feat": False,
        "resnetfpn": {"initial_dim": 128, "block_dims": [64, 128, 196, 256, 512, 1024]},
        "coarse": {"d_model": 256, "nhead": 8, "layer_names": 4, "attention": "linear"},
        "match_coarse": {
            "thr": 0.2, "border_rm": 2, "match_type": "dual_softmax", "dsmax_temperature": 0.1,
            "skh_iters": 3, "skh_init_bin_score": 1.0, "skh_prefilter": False,
            "train_coarse_percent": 0.2, "train_pad_num_gt_min": 200, "sparse_spvs": False,
        },
        "fine": {"d_model": 128, "nhead": 8, "layer_names": 1, "attention": "linear"},
    }


def load_reference_loftr(ckpt=None):
    """Build the reference nn.Module (eval mode, CPU fp32) with the shipped checkpoint."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_kornia_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from networks.loftr.loftr import LoFTR  # noqa: the unmodified reference

    model = LoFTR(default_loftr_config())
    ckpt = ckpt or os.path.join(REF_ROOT, "weights", "gim_loftr_50h.ckpt")
    sd = torch.load(ckpt, map_location="cpu", weights_only=True)
    sd = sd.get("state_dict", sd)
    missing = model.load_state_dict(dict(sd))
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    return model.eval()
