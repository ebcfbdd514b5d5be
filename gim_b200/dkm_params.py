"""Parameter containers of DKMv3 (state_dict keys identical to the reference's `RegressionMatcher` built by
networks/dkm/models/model_zoo/DKMv3.py:5-145) and a seeded, well-scaled random state_dict.

The trained `gim_dkm_100h.ckpt` is a git-LFS object that is absent from the reference tree, so parity
for this path is defined on a seeded random state_dict shared by the reference and by the CUDA path.  PyTorch's
default initialisation shrinks the signal by ~1/sqrt(3) per convolution (after ~40 layers every output is decided by
the last biases), which would make a parity test blind to upstream errors; `seeded_state_dict` therefore draws
variance-preserving weights (He for ReLU-fed layers), non-trivial BatchNorm statistics and small output heads, from a
`torch.Generator` on the CPU - reproducible on any machine with the same torch build, nothing to download or commit.
The modules below are parameter containers only (their `forward` is never called)."""
import math

import torch
import torch.nn as nn

REFINER_CFG = {  # scale: (in_dim, hidden_dim, displacement_emb_dim, local_corr_radius)   DKMv3.py:52-111
    "16": (2 * 512 + 128 + 15 ** 2, 2 * 512 + 128 + 15 ** 2, 128, 7),
    "8": (2 * 512 + 64 + 7 ** 2, 2 * 512 + 64 + 7 ** 2, 64, 3),
    "4": (2 * 256 + 32 + 5 ** 2, 2 * 256 + 32 + 5 ** 2, 32, 2),
    "2": (2 * 64 + 16, 128 + 16, 16, None),
    "1": (2 * 3 + 6, 24, 6, None),
}
HIDDEN_BLOCKS = 8
GP_DIM, DFN_DIM, FEAT_DIM = 256, 384, 256


def _dw_block(cin, cout):  # ConvRefiner.create_block (dkm.py:50-73) with dw=True
    return nn.Sequential(nn.Conv2d(cin, cout, 5, 1, 2, groups=cin), nn.BatchNorm2d(cout), nn.ReLU(), nn.Conv2d(cout, cout, 1))


class _ConvRefiner(nn.Module):  # dkm.py:11-48
    def __init__(self, in_dim, hidden, emb_dim):
        super().__init__()
        self.block1 = _dw_block(in_dim, hidden)
        self.hidden_blocks = nn.Sequential(*[_dw_block(hidden, hidden) for _ in range(HIDDEN_BLOCKS)])
        self.out_conv = nn.Conv2d(hidden, 3, 1)
        self.disp_emb = nn.Conv2d(2, emb_dim, 1)


class _RRB(nn.Module):  # dkm.py:173-202
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.bn = nn.BatchNorm2d(cout)
        self.conv3 = nn.Conv2d(cout, cout, 3, padding=1)


class _CAB(nn.Module):  # dkm.py:147-170
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 1)
        self.conv2 = nn.Conv2d(cout, cout, 1)


class _DFN(nn.Module):  # dkm.py:205-254, built at DKMv3.py:9-47
    def __init__(self):
        super().__init__()
        s = ("32", "16")
        self.feat_input_modules = nn.ModuleDict({k: nn.Conv2d(512, FEAT_DIM, 1) for k in s})
        self.pred_input_modules = nn.ModuleDict({k: nn.Identity() for k in s})
        self.rrb_d = nn.ModuleDict({k: _RRB(GP_DIM + FEAT_DIM, DFN_DIM) for k in s})
        self.cab = nn.ModuleDict({k: _CAB(2 * DFN_DIM, DFN_DIM) for k in s})
        self.rrb_u = nn.ModuleDict({k: _RRB(DFN_DIM, DFN_DIM) for k in s})
        self.terminal_module = nn.ModuleDict({k: nn.Conv2d(DFN_DIM, 3, 1) for k in s})


class _GP(nn.Module):  # dkm.py:257-280
    def __init__(self):
        super().__init__()
        self.pos_conv = nn.Conv2d(2, GP_DIM, 1)


class _Decoder(nn.Module):  # dkm.py:403-416
    def __init__(self):
        super().__init__()
        self.embedding_decoder = _DFN()
        self.gps = nn.ModuleDict({"32": _GP(), "16": _GP()})
        self.proj = nn.ModuleDict({"16": nn.Conv2d(1024, 512, 1), "32": nn.Conv2d(2048, 512, 1)})
        self.conv_refiner = nn.ModuleDict({k: _ConvRefiner(c[0], c[1], c[2]) for k, c in REFINER_CFG.items()})


class _Bottleneck(nn.Module):  # torchvision.models.resnet.Bottleneck (stride on conv2)
    def __init__(self, cin, planes, downsample):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(cin, planes, 1, bias=False), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, bias=False), nn.BatchNorm2d(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(cin, planes * 4, 1, bias=False), nn.BatchNorm2d(planes * 4))


class _ResNet50(nn.Module):  # torchvision resnet50 minus fc (encoders.py:30-45)
    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, bias=False), nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, nblk) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
            blocks = []
            for b in range(nblk):
                blocks.append(_Bottleneck(cin, planes, downsample=(b == 0)))
                cin = planes * 4
            setattr(self, f"layer{li}", nn.Sequential(*blocks))


class _Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.net = _ResNet50()


class DKMParams(nn.Module):
    """`encoder.net.*` + `decoder.*`: the reference RegressionMatcher's parameters and buffers."""

    def __init__(self):
        super().__init__()
        self.encoder = _Encoder()
        self.decoder = _Decoder()
        for p in self.parameters():
            p.requires_grad_(False)


def seeded_state_dict(seed=0):
    """A deterministic, well-scaled random state_dict with the reference's keys (see module docstring)."""
    g = torch.Generator().manual_seed(seed)
    model = DKMParams()
    out = {}
    for name, mod in model.named_modules():
        if isinstance(mod, nn.BatchNorm2d):
            n = mod.num_features
            lo, hi = (0.25, 0.5) if name.endswith("bn3") else (0.8, 1.2)  # damp the residual branch: no blow-up over 16 blocks
            out[name + ".weight"] = torch.rand(n, generator=g) * (hi - lo) + lo
            out[name + ".bias"] = torch.randn(n, generator=g) * 0.1
            out[name + ".running_mean"] = torch.randn(n, generator=g) * 0.1
            out[name + ".running_var"] = torch.rand(n, generator=g) * 0.5 + 0.75
            out[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        elif isinstance(mod, nn.Conv2d):
            shape = tuple(mod.weight.shape)
            fan_in = shape[1] * shape[2] * shape[3]
            gain = 2.0                                     # He: the layer feeds (BatchNorm +) ReLU
            if name.endswith("out_conv") or "terminal_module" in name:
                gain = 0.02                                # displacement / flow heads: a few pixels, coordinates stay in [-1, 1]
            elif name.endswith(".3") or "disp_emb" in name or "pos_conv" in name or ".proj." in name:
                gain = 1.0                                 # linear layers that are not followed by a ReLU
            out[name + ".weight"] = torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
            if mod.bias is not None:
                out[name + ".bias"] = torch.randn(shape[0], generator=g) * 0.05
    ref = model.state_dict()
    assert list(out) == list(ref) and all(out[k].shape == ref[k].shape for k in ref)
    return {k: v.to(ref[k].dtype) for k, v in out.items()}
