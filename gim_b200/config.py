"""Default gim_loftr configuration: `lower_config(get_cfg_defaults())['loftr']` of the reference
(networks/loftr/config.py:7-46, networks/loftr/misc.py:13-16) as a plain dict (yacs is not needed)."""
import copy

_DEFAULT = {
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


def get_default_config():
    return copy.deepcopy(_DEFAULT)


def lower_config(cfg):
    """Accepts a yacs CfgNode-like mapping or a dict; returns a lower-cased plain dict."""
    if not hasattr(cfg, "items"):
        return cfg
    return {str(k).lower(): lower_config(v) for k, v in cfg.items()}
