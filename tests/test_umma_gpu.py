"""tcgen05 split-fp16 GEMM (gim_b200/csrc/umma_gemm.cu) against a float64 reference and against the fp32
CUDA-core kernel, layer shape by layer shape (the shapes gim_loftr actually uses, plus tails)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT = {"none": 0, "relu": 1, "leaky": 2, "elu1": 3, "divs": 4}


def run_layer(B, H, W, C1, Cout, k, stride, C2=0, bn=False, residual=False, act="none", act1=None, split=1 << 30,
              div=1.0, mask=False, planes=True, seed=0, scale_in=1.0):
    from gim_b200 import _lib
    lib = _lib.load_test()
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, H, W, C1, generator=g) * scale_in).cuda()
    x2 = (torch.randn(B, H, W, C2, generator=g) * scale_in).cuda() if C2 else None
    Cin = C1 + C2
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).cuda()
    pad = k // 2
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    M = B * OH * OW
    sc = (torch.rand(Cout, generator=g) + 0.5).cuda() if bn else None
    bi = torch.randn(Cout, generator=g).cuda() if bn else None
    res = torch.randn(M, Cout, generator=g).cuda() if residual else None
    rm = (torch.rand(M, generator=g) > 0.3).to(torch.uint8).cuda() if mask else None
    out_u = torch.full((M, Cout), float("nan"), device="cuda")
    out_p = torch.full((M, Cout), float("nan"), device="cuda") if planes else None
    out_s = torch.full((M, Cout), float("nan"), device="cuda")
    ws = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    p = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
    a1 = ACT[act1] if act1 else ACT[act]
    rc = lib.gimb_test_conv(p(x), p(x2), B, H, W, C1, C2, p(w), Cout, k, stride, p(sc), p(bi), p(res), p(rm), ACT[act], a1,
                            split, div, p(out_u), p(out_p), p(out_s), ws.data_ptr(), ws.numel(),
                            torch.cuda.current_stream().cuda_stream)
    _lib.check_test(rc)
    torch.cuda.synchronize()
    # float64 reference
    xin = torch.cat([x, x2], -1) if C2 else x
    ref = F.conv2d(xin.permute(0, 3, 1, 2).double().cpu(), w.permute(0, 3, 1, 2).double().cpu(), stride=stride, padding=pad)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    if bn:
        ref = ref * sc.double().cpu() + bi.double().cpu()
    if residual:
        ref = ref + res.double().cpu()

    def actf(v, name):
        if name == "relu":
            return v.clamp(min=0)
        if name == "leaky":
            return torch.where(v > 0, v, 0.01 * v)
        if name == "elu1":
            return F.elu(v) + 1
        if name == "divs":
            return v / div
        return v
    if act1:
        ref = torch.cat([actf(ref[:, :split], act), actf(ref[:, split:], act1)], 1)
    else:
        ref = actf(ref, act)
    if mask:
        ref = ref * rm.double().cpu()[:, None]
    e_u = (out_u.double().cpu() - ref).abs().max().item()
    e_s = (out_s.double().cpu() - ref).abs().max().item()
    e_p = (out_p.double().cpu() - ref).abs().max().item() if planes else 0.0
    mag = ref.abs().max().item()
    return e_u, e_p, e_s, mag


SHAPES = [
    # B, H, W, C1, Cout, k, stride, kwargs
    (1, 300, 1, 256, 256, 1, 1, dict(planes=False)),                        # Linear with an M tail (fp32 out)
    (1, 300, 1, 256, 256, 1, 1, dict(residual=True)),                       # fp32 + planes + residual variant
    (1, 64, 64, 64, 64, 1, 1, dict(bn=True, act="relu")),                   # l1 conv1, N = 64
    (1, 4800, 1, 256, 512, 1, 1, dict(C2=256, act="relu")),                 # mlp.0 on cat[x, msg], two N tiles
    (2, 24, 40, 64, 64, 3, 1, dict(bn=True, act="relu")),                   # l1 conv2
    (1, 16, 32, 196, 196, 3, 1, dict(bn=True, act="leaky")),                # FPN 196-channel 3x3 (pitch 200, bn 208)
    (1, 16, 32, 196, 128, 3, 1, dict(planes=False)),                        # layer1_outconv2.3 (fp32 out)
    (2, 32, 48, 128, 128, 3, 2, dict(bn=True, act="relu")),                 # stride-2 3x3 (phase views)
    (2, 32, 48, 256, 512, 1, 2, dict(bn=True)),                             # stride-2 1x1 downsample
    (1, 60, 80, 256, 1024, 1, 1, dict(bn=True, residual=True, act="relu")), # l3 conv3 + identity, 60 rows (tile tail)
    (1, 60, 80, 1024, 256, 1, 1, dict(bn=True, act="relu")),                # K = 1024
    (1, 60, 80, 256, 256, 3, 1, dict(bn=True, act="relu")),                 # K = 2304
    (1, 1000, 1, 256, 512, 1, 1, dict(act="elu1", act1="divs", split=256, div=4800.0, mask=True, planes=False)),  # kv projection (fp32 out)
    (1, 777, 1, 128, 128, 1, 1, dict(act="elu1", mask=True, planes=False)),  # fine q projection (fp32 out)
    # >= 2 x 148 m-tiles: the CTA-pair path (TMA multicast of the weight tile), odd tile count -> one dummy tile
    (1, 38005, 1, 256, 256, 1, 1, dict(bn=True, act="relu", planes=False)),
    (3, 120, 160, 64, 64, 3, 1, dict(bn=True, act="relu", planes=False)),   # conv patches in CTA-pair mode (450 tiles)
]


@pytest.mark.parametrize("idx", range(len(SHAPES)))
def test_umma_matches_float64(idx):
    B, H, W, C1, Cout, k, stride, kw = SHAPES[idx]
    e_u, e_p, e_s, mag = run_layer(B, H, W, C1, Cout, k, stride, **kw)
    print(f"shape {SHAPES[idx]}: umma {e_u:.3e} planes {e_p:.3e} simt {e_s:.3e} |ref|max {mag:.3f}")
    tol = 4e-6 * max(mag, 1.0)   # fp32-class: a few ulp of the output magnitude
    assert e_s < tol, "fp32 CUDA-core kernel off"
    assert e_u < tol, "tcgen05 kernel (fp32 output) off"
    assert e_p < 2 * tol, "tcgen05 kernel (re-split planes) off"


def test_umma_large_values_do_not_overflow():
    # activations up to ~60 and weights up to ~4: 2^8 * hi stays far below the fp16 maximum
    e_u, e_p, e_s, mag = run_layer(1, 512, 1, 256, 256, 1, 1, scale_in=12.0, seed=3)
    print(f"large: umma {e_u:.3e} simt {e_s:.3e} mag {mag:.2f}")
    assert e_u < 4e-6 * mag
