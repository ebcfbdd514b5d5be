"""gim_b200 - B200 (sm_100a) native gim_loftr / gim_dkm dense image-pair matchers behind the reference's module API.

    from gim_b200 import LoFTR, load_default_weights
    model = LoFTR(get_default_config()); model.load_state_dict(load_default_weights()); model.eval().cuda()
    model(data)   # data['color0'|'color1'|'image0'|'image1'] -> data['mkpts0_f'|'mkpts1_f'|'mconf'|...]
"""
from .config import get_default_config, lower_config  # noqa: F401
from .weights import DEFAULT_WEIGHTS, load_gimw  # noqa: F401


def load_default_weights():
    """The shipped gim_loftr_50h weights as a state_dict (reference key names)."""
    return load_gimw(DEFAULT_WEIGHTS)


def __getattr__(name):
    if name == "LoFTR":
        from .loftr import LoFTR
        return LoFTR
    if name in ("DKMv3", "HlocDKM"):
        from . import dkm
        return getattr(dkm, name)
    raise AttributeError(name)
