"""TEST INFRASTRUCTURE - CPU oracle for the gim_loftr hot path.

A functional torch-CPU fp32 restatement of the reference forward pass
(`networks/loftr/loftr.py:43-91` and everything it calls).  It is NOT the product and
is never imported by `gim_b200/`; it exists so that `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg can check / time the CUDA path
against the reference algorithm on a box where /root/reference does not exist.

Parity pinning: `tests/test_oracle_golden.py` checks this file against golden vectors
generated from the UNMODIFIED reference (`oracle/make_golden.py`, run in the build
container through `oracle/ref_import.py`) - ids bit-exact, floats to <= 2e-5 - and, when
/root/reference is present, against the live reference module as well.

Weights: a flat dict {name: fp32 tensor} with the checkpoint's key names minus the
`model.` prefix (`networks/loftr/loftr.py:93-99`).
"""
import math

import torch
import torch.nn.functional as F

INF = 1e9  # networks/loftr/utils/coarse_matching.py:6


# ----------------------------------------------------------------------------- backbone
def _bn(x, w, pre, eps=1e-5):
    """nn.BatchNorm2d in eval mode (running statistics)."""
    return F.batch_norm(x, w[pre + ".running_mean"], w[pre + ".running_var"],
                        w[pre + ".weight"], w[pre + ".bias"], False, 0.0, eps)


def _bottleneck(x, w, pre, stride):
    """Bottleneck.forward, networks/loftr/backbone/resnet.py:106-126 (stride on the 3x3)."""
    out = F.relu(_bn(F.conv2d(x, w[pre + ".conv1.weight"]), w, pre + ".bn1"))
    out = F.relu(_bn(F.conv2d(out, w[pre + ".conv2.weight"], stride=stride, padding=1), w, pre + ".bn2"))
    out = _bn(F.conv2d(out, w[pre + ".conv3.weight"]), w, pre + ".bn3")
    if (pre + ".downsample.0.weight") in w:
        x = _bn(F.conv2d(x, w[pre + ".downsample.0.weight"], stride=stride), w, pre + ".downsample.1")
    return F.relu(out + x)


def resnet_trunk(x, w, pre="backbone.encode"):
    """ResNet._forward_impl, resnet.py:214-235: 7x7 s2 stem, NO max-pool, layers 1-3 only
    ([3, 4, 6] bottlenecks, resnet.py:272)."""
    x0 = F.relu(_bn(F.conv2d(x, w[pre + ".conv1.weight"], stride=2, padding=3), w, pre + ".bn1"))
    feats = []
    cur = x0
    for li, nblk in ((1, 3), (2, 4), (3, 6)):
        for b in range(nblk):
            stride = 2 if (li > 1 and b == 0) else 1
            cur = _bottleneck(cur, w, f"{pre}.layer{li}.{b}", stride)
        feats.append(cur)
    return feats  # x1 (1/2), x2 (1/4), x3 (1/8)


def _outconv2(x, w, pre):
    """conv3x3 -> BN -> LeakyReLU(0.01) -> conv3x3, resnet.py:278-289."""
    x = F.conv2d(x, w[pre + ".0.weight"], padding=1)
    x = F.leaky_relu(_bn(x, w, pre + ".1"), 0.01)
    return F.conv2d(x, w[pre + ".3.weight"], padding=1)


def backbone(x, w):
    """ResNetFPN_8_2.forward, resnet.py:306-329 -> (feat_c [B,256,H/8,W/8], feat_f [B,128,H/2,W/2])."""
    x1, x2, x3 = resnet_trunk(x, w)
    x3_out = F.conv2d(x3, w["backbone.layer3_outconv.weight"])
    x3_up = F.interpolate(x3_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x2_out = F.conv2d(x2, w["backbone.layer2_outconv.weight"])
    x2_out = _outconv2(x2_out + x3_up, w, "backbone.layer2_outconv2")
    x2_up = F.interpolate(x2_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x1_out = F.conv2d(x1, w["backbone.layer1_outconv.weight"])
    x1_out = _outconv2(x1_out + x2_up, w, "backbone.layer1_outconv2")
    return x3_out, x1_out


# ----------------------------------------------------------------------------- position encoding
def position_encoding(d_model, h, w_):
    """PositionEncodingSine with temp_bug_fix=False (loftr.py:22-24; position_encoding.py:22-37):
    `-math.log(10000.0) / d_model//2` parses as floor(-ln(1e4)/d_model) = -1.0, positions 1-based."""
    y_pos = torch.arange(1, h + 1, dtype=torch.float32).view(1, h, 1).expand(1, h, w_)
    x_pos = torch.arange(1, w_ + 1, dtype=torch.float32).view(1, 1, w_).expand(1, h, w_)
    coef = (-math.log(10000.0) / d_model) // 2
    div = torch.exp(torch.arange(0, d_model // 2, 2).float() * coef)[:, None, None]
    pe = torch.zeros(d_model, h, w_)
    pe[0::4] = torch.sin(x_pos * div)
    pe[1::4] = torch.cos(x_pos * div)
    pe[2::4] = torch.sin(y_pos * div)
    pe[3::4] = torch.cos(y_pos * div)
    return pe


# ----------------------------------------------------------------------------- transformer
def linear_attention(q, k, v, q_mask=None, kv_mask=None, eps=1e-6):
    """LinearAttention.forward, networks/loftr/submodules/attentions.py:20-47. q [N,L,H,D], k,v [N,S,H,D]."""
    Q = F.elu(q) + 1
    K = F.elu(k) + 1
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        v = v * kv_mask[:, :, None, None]
    S = v.size(1)
    v = v / S
    KV = torch.einsum("nshd,nshv->nhdv", K, v)
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return (torch.einsum("nlhd,nhdv,nlh->nlhv", Q, KV, Z) * S).contiguous()


def encoder_layer(x, src, w, pre, nhead, x_mask=None, src_mask=None):
    """LoFTREncoderLayer.forward, networks/loftr/submodules/transformer.py:35-58."""
    n, _, c = x.shape
    d = c // nhead
    q = F.linear(x, w[pre + ".q_proj.weight"]).view(n, -1, nhead, d)
    k = F.linear(src, w[pre + ".k_proj.weight"]).view(n, -1, nhead, d)
    v = F.linear(src, w[pre + ".v_proj.weight"]).view(n, -1, nhead, d)
    msg = linear_attention(q, k, v, x_mask, src_mask)
    msg = F.linear(msg.view(n, -1, c), w[pre + ".merge.weight"])
    msg = F.layer_norm(msg, (c,), w[pre + ".norm1.weight"], w[pre + ".norm1.bias"], 1e-5)
    msg = F.linear(torch.cat([x, msg], dim=2), w[pre + ".mlp.0.weight"])
    msg = F.linear(F.relu(msg), w[pre + ".mlp.2.weight"])
    msg = F.layer_norm(msg, (c,), w[pre + ".norm2.weight"], w[pre + ".norm2.bias"], 1e-5)
    return x + msg


def local_feature_transformer(f0, f1, w, pre, n_pairs_of_layers, nhead, m0=None, m1=None):
    """LocalFeatureTransformer.forward, transformer.py:80-101: (self, cross) x n; the second
    cross call sees the already-updated feat0 (:96-97)."""
    for i in range(n_pairs_of_layers):
        ps, pc = f"{pre}.layers.{2 * i}", f"{pre}.layers.{2 * i + 1}"
        f0 = encoder_layer(f0, f0, w, ps, nhead, m0, m0)
        f1 = encoder_layer(f1, f1, w, ps, nhead, m1, m1)
        f0 = encoder_layer(f0, f1, w, pc, nhead, m0, m1)
        f1 = encoder_layer(f1, f0, w, pc, nhead, m1, m0)
    return f0, f1


# ----------------------------------------------------------------------------- coarse matching
def dual_softmax_conf(fc0, fc1, temperature=0.1, m0=None, m1=None):
    """CoarseMatching.forward dual-softmax branch, networks/loftr/utils/coarse_matching.py:106-118."""
    c = fc0.shape[-1]
    fc0, fc1 = fc0 / c ** 0.5, fc1 / c ** 0.5
    sim = torch.einsum("nlc,nsc->nls", fc0, fc1) / temperature
    if m0 is not None:
        sim.masked_fill_(~(m0[..., None] * m1[:, None]).bool(), -INF)
    return F.softmax(sim, 1) * F.softmax(sim, 2)


def coarse_select(conf, hw0_c, hw1_c, thr=0.2, border=2, mask0=None, mask1=None):
    """CoarseMatching.get_coarse_match (eval mode), coarse_matching.py:150-195, with
    mask_border (:9-26) / mask_border_with_padding (:29-44).  Returns b, i, j (int64) and mconf."""
    n = conf.shape[0]
    h0, w0 = hw0_c
    h1, w1 = hw1_c
    m = (conf > thr).view(n, h0, w0, h1, w1).clone()
    if border > 0:
        m[:, :border] = False
        m[:, :, :border] = False
        m[:, :, :, :border] = False
        m[:, :, :, :, :border] = False
        if mask0 is None:
            m[:, -border:] = False
            m[:, :, -border:] = False
            m[:, :, :, -border:] = False
            m[:, :, :, :, -border:] = False
        else:
            h0s, w0s = mask0.sum(1).max(-1)[0].int(), mask0.sum(-1).max(-1)[0].int()
            h1s, w1s = mask1.sum(1).max(-1)[0].int(), mask1.sum(-1).max(-1)[0].int()
            for b in range(n):
                m[b, h0s[b] - border:] = False
                m[b, :, w0s[b] - border:] = False
                m[b, :, :, h1s[b] - border:] = False
                m[b, :, :, :, w1s[b] - border:] = False
    m = m.view(n, h0 * w0, h1 * w1)
    m = m * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j = m.max(dim=2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j[b_ids, i_ids]
    return b_ids, i_ids, j_ids, conf[b_ids, i_ids, j_ids]


# ----------------------------------------------------------------------------- fine level
def fine_windows(feat_f, b_ids, ids, hw_c, W=5):
    """FinePreprocess.forward, networks/loftr/submodules/fine_preprocess.py:29-47, without
    materialising the unfold: window m covers rows stride*r-W//2 .. +W//2 (zero padded) of the
    fine map, token order row-major, returns [M, W*W, C]."""
    n, c, hf, wf = feat_f.shape
    stride = hf // hw_c[0]
    r, q = ids // hw_c[1], ids % hw_c[1]
    pad = W // 2
    fp = F.pad(feat_f, (pad, pad, pad, pad))
    dy, dx = torch.meshgrid(torch.arange(W), torch.arange(W), indexing="ij")
    yy = (r * stride)[:, None] + dy.reshape(1, -1)
    xx = (q * stride)[:, None] + dx.reshape(1, -1)
    return fp[b_ids[:, None], :, yy, xx]  # [M, WW, C]


def fine_matching(f0, f1):
    """FineMatching.forward, networks/loftr/utils/fine_matching.py:43-57 (kornia
    spatial_expectation2d / create_meshgrid restated: grid = linspace(-1,1,W)^2, (x, y))."""
    m, ww, c = f0.shape
    W = int(math.sqrt(ww))
    sim = torch.einsum("mc,mrc->mr", f0[:, ww // 2, :], f1)
    heat = torch.softmax(sim / c ** 0.5, dim=1)
    lin = torch.linspace(-1, 1, W)
    gy, gx = torch.meshgrid(lin, lin, indexing="ij")
    grid = torch.stack([gx.reshape(-1), gy.reshape(-1)], -1)  # [WW, 2]
    coords = torch.stack([(heat * grid[:, 0]).sum(-1), (heat * grid[:, 1]).sum(-1)], -1)
    var = torch.sum(grid[None] ** 2 * heat[..., None], dim=1) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    return coords, torch.cat([coords, std[:, None]], -1)


# ----------------------------------------------------------------------------- whole forward
@torch.no_grad()
def loftr_forward(w, data, cfg=None, return_intermediates=False):
    """LoFTR.forward (networks/loftr/loftr.py:43-91), eval mode.  `data` needs color0/color1
    ([N,3,H,W] fp32 in [0,1], H,W multiples of 8); optional mask0/mask1 [N,H/8,W/8] and
    scale0/scale1 [N,2].  Returns a dict with the reference's output keys."""
    thr, border, temp, W = 0.2, 2, 0.1, 5
    if cfg is not None:
        thr, border = cfg["match_coarse"]["thr"], cfg["match_coarse"]["border_rm"]
        temp, W = cfg["match_coarse"]["dsmax_temperature"], cfg["fine_window_size"]
    c0, c1 = data["color0"], data["color1"]
    n = c0.shape[0]
    out = {"bs": n, "hw0_i": tuple(c0.shape[2:]), "hw1_i": tuple(c1.shape[2:])}
    if c0.shape == c1.shape:  # loftr.py:58-61
        fc, ff = backbone(torch.cat([c0, c1], 0), w)
        (fc0, fc1), (ff0, ff1) = fc.split(n), ff.split(n)
    else:
        (fc0, ff0), (fc1, ff1) = backbone(c0, w), backbone(c1, w)
    hw0_c, hw1_c = tuple(fc0.shape[2:]), tuple(fc1.shape[2:])
    hw0_f, hw1_f = tuple(ff0.shape[2:]), tuple(ff1.shape[2:])
    out.update(hw0_c=hw0_c, hw1_c=hw1_c, hw0_f=hw0_f, hw1_f=hw1_f)
    inter = {"feat_c0_backbone": fc0, "feat_c1_backbone": fc1, "feat_f0": ff0, "feat_f1": ff1}

    d = fc0.shape[1]
    t0 = (fc0 + position_encoding(d, *hw0_c)[None]).flatten(2).transpose(1, 2)
    t1 = (fc1 + position_encoding(d, *hw1_c)[None]).flatten(2).transpose(1, 2)
    m0 = m1 = None
    if "mask0" in data:
        m0, m1 = data["mask0"].flatten(-2), data["mask1"].flatten(-2)
    t0, t1 = local_feature_transformer(t0, t1, w, "loftr_coarse", 4, 8, m0, m1)
    inter.update(feat_c0=t0, feat_c1=t1)

    conf = dual_softmax_conf(t0, t1, temp, m0, m1)
    b_ids, i_ids, j_ids, mconf = coarse_select(conf, hw0_c, hw1_c, thr, border,
                                               data.get("mask0"), data.get("mask1"))
    scale = out["hw0_i"][0] / hw0_c[0]  # coarse_matching.py:240-247
    s0 = scale * data["scale0"][b_ids] if "scale0" in data else scale
    s1 = scale * data["scale1"][b_ids] if "scale1" in data else scale
    mk0_c = torch.stack([i_ids % hw0_c[1], i_ids // hw0_c[1]], 1) * s0
    mk1_c = torch.stack([j_ids % hw1_c[1], j_ids // hw1_c[1]], 1) * s1
    out.update(b_ids=b_ids, i_ids=i_ids, j_ids=j_ids, m_bids=b_ids, mconf=mconf,
               mkpts0_c=mk0_c.float(), mkpts1_c=mk1_c.float(), gt_mask=mconf == 0, W=W)
    if return_intermediates:
        inter["conf_matrix"] = conf

    if b_ids.numel() == 0:  # fine_matching.py:33-41
        out.update(expec_f=torch.empty(0, 3), mkpts0_f=out["mkpts0_c"], mkpts1_f=out["mkpts1_c"])
    else:
        w0 = fine_windows(ff0, b_ids, i_ids, hw0_c, W)
        w1 = fine_windows(ff1, b_ids, j_ids, hw1_c, W)
        w0, w1 = local_feature_transformer(w0, w1, w, "loftr_fine", 1, 8)
        coords, expec = fine_matching(w0, w1)
        fscale = out["hw0_i"][0] / hw0_f[0]  # fine_matching.py:63-69
        fs1 = fscale * data["scale1"][b_ids] if "scale0" in data else fscale
        out.update(expec_f=expec, mkpts0_f=out["mkpts0_c"],
                   mkpts1_f=out["mkpts1_c"] + coords * (W // 2) * fs1)
        inter.update(fine_win0=w0, fine_win1=w1)
    if return_intermediates:
        out["_inter"] = inter
    return out
