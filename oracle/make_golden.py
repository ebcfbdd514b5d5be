"""TEST INFRASTRUCTURE: generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden
Every case stores its inputs as uint8 (color = u8.float()/255) plus the reference's outputs
(`networks/loftr/loftr.py:43-91`, CPU fp32, torch.get_num_threads() threads).
"""
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gim_b200 import synth  # noqa: E402
from oracle.ref_import import REF_ROOT, load_reference_loftr  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
OUT_KEYS = ["b_ids", "i_ids", "j_ids", "m_bids", "mconf", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f", "expec_f"]


def u8(img):  # float [..,3,h,w] on the u8 grid -> uint8
    return torch.round(img * 255.0).to(torch.uint8).numpy()


def f32(a):
    return torch.from_numpy(a).float() / 255.0


def demo_image(name, h, w):
    im = cv2.imread(os.path.join(REF_ROOT, "assets", "demo", name + ".png"))[:, :, ::-1]
    im = cv2.resize(im, (w, h), interpolation=cv2.INTER_AREA)
    return torch.from_numpy(np.ascontiguousarray(im)).permute(2, 0, 1)[None].float() / 255.0


def run(model, data, hooks=False):
    inter = {}
    handles = []
    if hooks:
        handles.append(model.backbone.register_forward_hook(
            lambda m, i, o: inter.update(feat_c_backbone=o[0].numpy().copy(), feat_f=o[1].numpy().copy())))
        handles.append(model.loftr_coarse.register_forward_hook(
            lambda m, i, o: inter.update(feat_c0=o[0].numpy().copy(), feat_c1=o[1].numpy().copy())))
        handles.append(model.loftr_fine.register_forward_hook(
            lambda m, i, o: inter.update(fine_win0=o[0].numpy().copy(), fine_win1=o[1].numpy().copy())))
    t = time.time()
    with torch.no_grad():
        model(data)
    dt = time.time() - t
    for h in handles:
        h.remove()
    out = {k: data[k].numpy() for k in OUT_KEYS}
    out.update({"inter_" + k: v for k, v in inter.items()})
    return out, dt


def save(name, inputs, out, dt):
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **inputs, **out)
    print(f"{name}: M={len(out['b_ids'])} ref_forward={dt:.2f}s", flush=True)


def main():
    torch.manual_seed(0)
    model = load_reference_loftr()
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    if want("tiny_64x96"):  # stage intermediates
        c0, c1 = synth.make_pairs(1, 64, 96, first=0)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data, hooks=True)
        save("tiny_64x96", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("small_b2_240x320"):  # batch of 2, ordering by (b, i)
        c0, c1 = synth.make_pairs(2, 240, 320, first=1)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data, hooks=True)
        for k in list(out):
            if k.startswith("inter_") and k not in ("inter_feat_c0", "inter_feat_c1"):
                del out[k]
        save("small_b2_240x320", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("noise_96x128"):  # uniform noise: a handful of spurious matches
        g = torch.Generator().manual_seed(7)
        c0 = torch.round(torch.rand(1, 3, 96, 128, generator=g) * 255) / 255
        c1 = torch.round(torch.rand(1, 3, 96, 128, generator=g) * 255) / 255
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("noise_96x128", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("nomatch_flat_96x128"):  # constant images: M == 0 shortcut (fine_matching.py:33-41)
        c0 = torch.full((1, 3, 96, 128), 128 / 255.0)
        c1 = torch.full((1, 3, 96, 128), 64 / 255.0)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("nomatch_flat_96x128", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("diffsize_256x320_320x256"):  # L != S, two backbone calls (loftr.py:62-63)
        c0 = demo_image("b1", 256, 320)
        c1 = demo_image("b2", 320, 256)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("diffsize_256x320_320x256", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("masked_scaled_b2_256x320"):  # ZEB-style padding masks + scale0/scale1
        c0, c1 = synth.make_pairs(2, 256, 320, first=3)
        valid = [(208, 320, 256, 264), (256, 240, 224, 320)]  # (h0, w0, h1, w1) valid extents, multiples of 8
        m0 = torch.zeros(2, 32, 40, dtype=torch.bool)
        m1 = torch.zeros(2, 32, 40, dtype=torch.bool)
        for b, (h0, w0, h1, w1) in enumerate(valid):
            c0[b, :, h0:, :] = 0
            c0[b, :, :, w0:] = 0
            c1[b, :, h1:, :] = 0
            c1[b, :, :, w1:] = 0
            m0[b, :h0 // 8, :w0 // 8] = True
            m1[b, :h1 // 8, :w1 // 8] = True
        s0 = torch.tensor([[1.5, 1.25], [2.0, 2.0]])
        s1 = torch.tensor([[1.0, 1.0], [1.75, 2.5]])
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1, mask0=m0, mask1=m1, scale0=s0, scale1=s1)
        out, dt = run(model, data)
        save("masked_scaled_b2_256x320",
             dict(color0_u8=u8(c0), color1_u8=u8(c1), mask0=m0.numpy(), mask1=m1.numpy(),
                  scale0=s0.numpy(), scale1=s1.numpy()), out, dt)

    if want("demo_a_480x640"):  # real pair a1 <-> a2 at the headline resolution
        c0, c1 = demo_image("a1", 480, 640), demo_image("a2", 480, 640)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("demo_a_480x640", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    if want("synth_b2_480x640"):  # two bench-workload pairs (bench.py config), batch 2
        c0, c1 = synth.make_pairs(2, 480, 640, first=0)
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("synth_b2_480x640", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    # ---- round-2 cases: the configurations BASELINE.json names beyond 480x640 (inputs are regenerated from the
    # recipe at test time, see tests/goldens.py::RECIPES; only the demo pair stores its pixels)
    if want("demo_a_1000x1000"):  # config 1: a1 <-> a2 after demo.py:151-177 (floor to x8 -> 1000x1000), on the u8 grid
        import torchvision.transforms.functional as TF
        cs = []
        for nm in ("a1", "a2"):
            im = cv2.imread(os.path.join(REF_ROOT, "assets", "demo", nm + ".png"))[:, :, ::-1]
            t = torch.from_numpy(np.ascontiguousarray(im).astype(np.float32).transpose(2, 0, 1) / 255.0).float()
            size_new = tuple(int(x // 8 * 8) for x in t.shape[-2:])
            t = TF.resize(t, size=size_new)
            cs.append((torch.round(t.clamp(0, 1) * 255.0) / 255.0)[None])
        c0, c1 = cs
        data = dict(color0=c0, color1=c1, image0=c0, image1=c1)
        out, dt = run(model, data)
        save("demo_a_1000x1000", dict(color0_u8=u8(c0), color1_u8=u8(c1)), out, dt)

    from tests.goldens import RECIPES, build_recipe  # noqa: E402
    for name in RECIPES:
        if want(name):
            data = build_recipe(name)
            out, dt = run(model, data)
            save(name, {}, out, dt)


if __name__ == "__main__":
    main()
