"""TEST INFRASTRUCTURE: generate tests/golden/dkm_*.npz from the UNMODIFIED reference DKMv3.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_dkm
The trained gim_dkm checkpoint is absent from the reference tree (git-LFS), so both sides load the seeded state_dict of
gim_b200/dkm_params.py; parity is defined on match()'s dense (warp, certainty).  Inputs are
stored as uint8 (image = u8 / 255)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gim_b200 import synth  # noqa: E402
from gim_b200.dkm_params import seeded_state_dict  # noqa: E402
from oracle.ref_import import REF_ROOT  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {  # name: (input h, w, h_resized, w_resized, upsample_res, synth pair index)
    "dkm_64x96_up128x192": (80, 112, 64, 96, (128, 192), 3),
    "dkm_96x128_up192x256": (120, 160, 96, 128, (192, 256), 5),
    # sizes that are NOT multiples of 32 (the ZEB harness runs 660 x 880, trainer/lightning.py:33-34): 88 -> 44 -> 22 -> 11 -> 6 -> 3
    # and 120 -> 60 -> 30 -> 15 -> 8 -> 4, i.e. stride-2 convolutions on odd maps; second pass 180 x 244 -> 90 x 122 -> 45 x 61 -> 23 x 31
    "dkm_odd_88x120_up180x244": (100, 140, 88, 120, (180, 244), 9),
    # big enough for the tensor-core Gram path of the GP (N = 14 * 18 = 252 tokens at 1/16) and several Cholesky blocks
    "dkm_224x288_up320x416": (240, 320, 224, 288, (320, 416), 7),
    # BASELINE config 3 geometry (672x896 -> 1152x1536 second pass): outputs stored on a stride-8 grid (the full tensors
    # are 70 MB); GP with 588 / 2352 tokens, all refiner shapes at their real sizes
    "dkm_672x896_up1152x1536_s8": (480, 640, 672, 896, (1152, 1536), 11),
}


def load_reference_dkm(h, w, upsample_res, seed=0):
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from networks.dkm.models.model_zoo.DKMv3 import DKMv3
    m = DKMv3(None, h, w, upsample_preds=True)
    m.load_state_dict(seeded_state_dict(seed))
    m.upsample_res = upsample_res
    return m.eval()


def case_images(name):
    ih, iw, _, _, _, idx = CASES[name]
    a, b = synth.make_pairs(1, ih, iw, first=idx)
    a = torch.round(a * 255).to(torch.uint8)
    b = torch.round(b * 255).to(torch.uint8)
    b[:, :, : ih // 6, : iw // 5] = 0  # a black corner: exercises the black-pixel mask (models/dkm.py:726-731)
    return a, b


def main():
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, (ih, iw, h, w, up, _) in CASES.items():
        if only and name not in only:
            continue
        a, b = case_images(name)
        m = load_reference_dkm(h, w, up)
        warp, cert = m.match(a.float() / 255, b.float() / 255)
        if name.endswith("_s8"):
            warp, cert = warp[::8, ::8].contiguous(), cert[::8, ::8].contiguous()
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), im0_u8=a.numpy(), im1_u8=b.numpy(), h=h, w=w, up=np.array(up),
                            warp=warp.numpy(), certainty=cert.numpy())
        print(name, tuple(warp.shape), tuple(cert.shape), "certainty mean", float(cert.mean()), "nonzero", float((cert > 0).float().mean()))


if __name__ == "__main__":
    main()
