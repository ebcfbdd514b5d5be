"""TEST INFRASTRUCTURE - CPU restatement of the gim_dkm (DKMv3) dense matcher, fp32 torch, state_dict driven.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(gim_b200/dkm.py -> libgimb200.so) never does.  Every function cites the reference lines it restates
(paths relative to /root/reference/networks/dkm).

PINNED: tests/test_dkm_oracle.py checks it against tests/golden/dkm_*.npz, which oracle/make_golden_dkm.py produced
from the UNMODIFIED reference `DKMv3(None, h, w)` with the seeded state_dict of gim_b200/dkm_params.py, and - where
/root/reference exists - against the live reference module."""
import math

import torch
import torch.nn.functional as F


def _conv(x, sd, pre, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[pre + ".weight"], sd.get(pre + ".bias"), stride=stride, padding=padding, groups=groups)


def _bn(x, sd, pre):
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"], sd[pre + ".weight"], sd[pre + ".bias"],
                        False, 0.0, 1e-5)


def _grid(h, w):
    """torch.meshgrid(linspace(-1+1/h, 1-1/h, h), linspace(-1+1/w, 1-1/w, w)) stacked (x, y): [2, h, w]
    (models/dkm.py:88-95, 324-337, 439-451)."""
    ys = torch.linspace(-1 + 1 / h, 1 - 1 / h, h)
    xs = torch.linspace(-1 + 1 / w, 1 - 1 / w, w)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gx, gy))


def _bottleneck(x, sd, pre, stride):  # torchvision.models.resnet.Bottleneck (v1.5: the stride sits on conv2)
    out = F.relu(_bn(_conv(x, sd, pre + ".conv1"), sd, pre + ".bn1"))
    out = F.relu(_bn(_conv(out, sd, pre + ".conv2", stride=stride, padding=1), sd, pre + ".bn2"))
    out = _bn(_conv(out, sd, pre + ".conv3"), sd, pre + ".bn3")
    idn = x
    if (pre + ".downsample.0.weight") in sd:
        idn = _bn(_conv(x, sd, pre + ".downsample.0", stride=stride), sd, pre + ".downsample.1")
    return F.relu(out + idn)


def encoder(x, sd, upto=32):
    """ResNet50.forward (models/encoders.py:46-62): {1: image, 2, 4, 8, 16, 32}."""
    p = "encoder.net."
    feats = {1: x}
    x = F.relu(_bn(_conv(x, sd, p + "conv1", stride=2, padding=3), sd, p + "bn1"))
    feats[2] = x
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (nblk, scale) in enumerate(((3, 4), (4, 8), (6, 16), (3, 32)), start=1):
        if scale > upto:
            break
        for b in range(nblk):
            x = _bottleneck(x, sd, f"{p}layer{li}.{b}", 2 if (b == 0 and li > 1) else 1)
        feats[scale] = x
    return feats


def cos_kernel(x, y, T=0.2, eps=1e-6):
    """CosKernel.__call__ (models/dkm.py:135-144)."""
    c = torch.einsum("bnd,bmd->bnm", x, y) / (x.norm(dim=-1)[..., None] * y.norm(dim=-1)[:, None] + eps)
    return ((c - 1.0) / T).exp()


def gp(x, y, sd, pre, sigma_noise=0.1):
    """GP.forward with no_cov=True, basis='fourier' (models/dkm.py:340-370; get_pos_enc :324-338)."""
    b, c, h1, w1 = x.shape
    _, _, h2, w2 = y.shape
    coords = _grid(h2, w2)[None].expand(b, 2, h2, w2)
    f = torch.cos(8 * math.pi * _conv(coords, sd, pre + ".pos_conv"))
    xr, yr, fr = (t.flatten(2).transpose(1, 2) for t in (x, y, f))
    K_yy = cos_kernel(yr, yr)
    K_xy = cos_kernel(xr, yr)
    A = K_yy + sigma_noise * torch.eye(h2 * w2)[None]
    if h2 * w2 > 2000:
        # REFERENCE QUIRK, kept because parity is defined against the unmodified reference (dkm.py:352-356): above 2000
        # tokens it inverts "one matrix at a time" with `sigma_noise[k:k+1]`, but sigma_noise has batch size 1, so for
        # k >= 1 the slice is empty, the sum broadcasts to an empty batch and only the FIRST matrix is inverted; the matmul
        # below then broadcasts that one inverse over the whole batch (the support->query half uses the query->support
        # half's K_yy^-1).  672x896 inputs have 2352 tokens at 1/16 and take this branch.
        K_yy_inv = torch.linalg.inv(A[0:1])
    else:
        # one matrix at a time: the batched routine runs the same LAPACK factorisation per matrix, and its MKL build fails
        # on some hosts ("SLASWP parameter 6")
        K_yy_inv = torch.cat([torch.linalg.inv(A[k:k + 1]) for k in range(b)])
    mu = K_xy.matmul(K_yy_inv.matmul(fr))
    return mu.transpose(1, 2).reshape(b, -1, h1, w1)


def rrb(x, sd, pre):
    """RRB.forward (models/dkm.py:196-202)."""
    x = _conv(x, sd, pre + ".conv1")
    res = F.relu(_bn(_conv(x, sd, pre + ".conv2", padding=1), sd, pre + ".bn"))
    res = _conv(res, sd, pre + ".conv3", padding=1)
    return F.relu(x + res)


def cab(x1, x2, sd, pre):
    """CAB.forward (models/dkm.py:160-170): x1 = old context, x2 = new embeddings."""
    x = torch.cat([x1, x2], dim=1).mean(dim=(2, 3), keepdim=True)
    x = torch.sigmoid(_conv(F.relu(_conv(x, sd, pre + ".conv1")), sd, pre + ".conv2"))
    return x * x2 + x1


def dfn(emb, feats, context, sd, key):
    """DFN.forward (models/dkm.py:245-254)."""
    p = "decoder.embedding_decoder."
    feats = _conv(feats, sd, f"{p}feat_input_modules.{key}")
    emb = rrb(torch.cat([feats, emb], dim=1), sd, f"{p}rrb_d.{key}")
    context = cab(context, emb, sd, f"{p}cab.{key}")
    context = rrb(context, sd, f"{p}rrb_u.{key}")
    preds = _conv(context, sd, f"{p}terminal_module.{key}")
    return preds[:, -2:], preds[:, :-2], context


def local_correlation(f0, f1, r, flow):
    """utils/local_correlation.py:5-40 (corr_in_other: a (2r+1)^2 window around the flow target in the other image)."""
    b, c, h, w = f0.shape
    coords = flow.permute(0, 2, 3, 1)
    wy, wx = torch.meshgrid(torch.linspace(-2 * r / h, 2 * r / h, 2 * r + 1), torch.linspace(-2 * r / w, 2 * r / w, 2 * r + 1),
                            indexing="ij")
    win = torch.stack((wx, wy), dim=-1)[None].expand(b, 2 * r + 1, 2 * r + 1, 2).reshape(b, (2 * r + 1) ** 2, 2)
    coords = (coords[:, :, :, None] + win[:, None, None]).reshape(b, h, w * (2 * r + 1) ** 2, 2)
    wf = F.grid_sample(f1, coords, padding_mode="zeros", align_corners=False)[..., None].reshape(b, c, h, w, (2 * r + 1) ** 2)
    return torch.einsum("bchw, bchwk -> bkhw", f0, wf) / (c ** 0.5)


def conv_refiner(x, y, flow, sd, key, radius, taps=None):
    """ConvRefiner.forward (models/dkm.py:75-123) as configured by model_zoo/DKMv3.py:52-111
    (displacement_emb='linear', corr_in_other=True, depthwise 5x5 blocks)."""
    p = f"decoder.conv_refiner.{key}."
    b, c, hs, ws = x.shape
    x_hat = F.grid_sample(y, flow.permute(0, 2, 3, 1), align_corners=False)
    disp = flow - _grid(hs, ws)[None]
    emb = _conv(disp, sd, p + "disp_emb")
    parts = [x, x_hat, emb]
    if radius:
        parts.append(local_correlation(x, y, radius, flow))
    d = torch.cat(parts, dim=1)
    if taps is not None and f"refiner_in{key}" not in taps:
        taps[f"refiner_in{key}"] = d

    def block(d, pre):
        d = _conv(d, sd, pre + ".0", padding=2, groups=d.shape[1])  # depthwise (dw=True), out_dim a multiple of in_dim
        d = F.relu(_bn(d, sd, pre + ".1"))
        if taps is not None and f"refiner_dw{key}" not in taps:
            taps[f"refiner_dw{key}"] = d
        return _conv(d, sd, pre + ".3")

    d = block(d, p + "block1")
    if taps is not None and f"refiner_pw{key}" not in taps:
        taps[f"refiner_pw{key}"] = d
    for i in range(8):
        d = block(d, f"{p}hidden_blocks.{i}")
    if taps is not None and f"refiner_out{key}" not in taps:
        taps[f"refiner_out{key}"] = d
    d = _conv(d, sd, p + "out_conv")
    return d[:, :-2], d[:, -2:]


REFINER_RADIUS = {"16": 7, "8": 3, "4": 2, "2": None, "1": None}


def decoder(f1, f2, sd, upsample=False, dense_flow=None, dense_certainty=None, taps=None):
    """Decoder.forward (models/dkm.py:454-534)."""
    scales = ["8", "4", "2", "1"] if upsample else ["32", "16", "8", "4", "2", "1"]
    sizes = {s: f1[s].shape[-2:] for s in f1}
    h, w = sizes[1]
    b = f1[1].shape[0]
    coarsest = int(scales[0])
    old = torch.zeros(b, 384, *sizes[coarsest])
    out = {}
    if not upsample:
        dense_flow = _grid(*sizes[coarsest])[None].expand(b, 2, *sizes[coarsest])
        dense_certainty = 0.0
    else:
        dense_flow = F.interpolate(dense_flow, size=sizes[coarsest], align_corners=False, mode="bilinear")
        dense_certainty = F.interpolate(dense_certainty, size=sizes[coarsest], align_corners=False, mode="bilinear")
    for s in scales:
        ins = int(s)
        a, c = f1[ins], f2[ins]
        if s in ("16", "32"):
            a, c = _conv(a, sd, f"decoder.proj.{s}"), _conv(c, sd, f"decoder.proj.{s}")
            old = F.interpolate(old, size=sizes[ins], mode="bilinear", align_corners=False)
            new = gp(a, c, sd, f"decoder.gps.{s}")
            if taps is not None:
                taps[f"gp{s}"] = new
            dense_flow, dense_certainty, old = dfn(new, a, old, sd, s)
            if taps is not None:
                taps[f"dfn_flow{s}"] = dense_flow
        dc, disp = conv_refiner(a, c, dense_flow, sd, s, REFINER_RADIUS[s], taps) if s != "32" else (None, None)
        if s != "32":
            dense_flow = torch.stack((dense_flow[:, 0] + ins * disp[:, 0] / (4 * w), dense_flow[:, 1] + ins * disp[:, 1] / (4 * h)), dim=1)
            dense_certainty = dense_certainty + dc
        out[ins] = {"dense_flow": dense_flow, "dense_certainty": dense_certainty}
        if taps is not None:
            taps[f"flow{s}{'u' if upsample else ''}"] = dense_flow
            taps[f"cert{s}{'u' if upsample else ''}"] = dense_certainty
        if s != "1":
            dense_flow = F.interpolate(dense_flow, size=sizes[ins // 2], align_corners=False, mode="bilinear")
            dense_certainty = F.interpolate(dense_certainty, size=sizes[ins // 2], align_corners=False, mode="bilinear")
    return out


def forward_symmetric(q, s, sd, upsample=False, corresps=None, taps=None):
    """RegressionMatcher.forward_symmetric (models/dkm.py:640-650): batch = cat(query, support); the support pyramid is
    the same features with the two halves swapped."""
    pyr = encoder(torch.cat((q, s)), sd, upto=8 if upsample else 32)
    if taps is not None and not upsample:
        for k in (2, 4, 8, 16, 32):
            taps[f"enc{k}"] = pyr[k]
    swp = {k: torch.cat((v.chunk(2)[1], v.chunk(2)[0])) for k, v in pyr.items()}
    kw = {} if corresps is None else corresps
    return decoder(pyr, swp, sd, upsample=upsample, taps=taps, **kw)


def match(sd, im1, im2, h, w, upsample_res, upsample_preds=True, taps=None):
    """RegressionMatcher.match with symmetric=True, batched=False (models/dkm.py:655-752).
    Returns warp [H, 2W, 4] and certainty [H, 2W]."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    with torch.no_grad():
        q = F.interpolate(im1, size=(h, w), mode="bilinear", align_corners=False)
        s = F.interpolate(im2, size=(h, w), mode="bilinear", align_corners=False)
        cor = forward_symmetric(q, s, sd, taps=taps)
        hs, ws = upsample_res if upsample_preds else (h, w)
        low = F.interpolate(cor[16]["dense_certainty"], size=(hs, ws), align_corners=False, mode="bilinear")
        low = 0.5 * low * (low < 0)
        if upsample_preds:
            q = F.interpolate(im1, size=(hs, ws), mode="bilinear", align_corners=False)
            s = F.interpolate(im2, size=(hs, ws), mode="bilinear", align_corners=False)
            cor = forward_symmetric(q, s, sd, upsample=True, corresps=cor[1], taps=taps)
        q2s = cor[1]["dense_flow"].permute(0, 2, 3, 1)
        cert = (cor[1]["dense_certainty"] - low).sigmoid()
        qc = _grid(hs, ws)[None].expand(1, 2, hs, ws).permute(0, 2, 3, 1)
        wrong = (q2s.abs() > 1).sum(dim=-1) > 0
        cert[wrong[:, None]] = 0
        b1 = (im1[0, 0] < 0.03125) & (im1[0, 1] < 0.03125) & (im1[0, 2] < 0.03125)
        b2 = (im2[0, 0] < 0.03125) & (im2[0, 1] < 0.03125) & (im2[0, 2] < 0.03125)
        b1 = F.interpolate(b1.float()[None, None], size=(hs, ws), mode="nearest").bool()
        b2 = F.interpolate(b2.float()[None, None], size=(hs, ws), mode="nearest").bool()
        cert[torch.cat((b1, b2), dim=0)] = 0
        q2s = torch.clamp(q2s, -1, 1)
        qts, stq = q2s.chunk(2)
        warp = torch.cat((torch.cat((qc, qts), dim=-1), torch.cat((stq, qc), dim=-1)), dim=2)
        cert = torch.cat(cert.chunk(2), dim=3)[:, 0]
    return warp[0], cert[0]
