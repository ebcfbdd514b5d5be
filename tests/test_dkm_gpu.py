"""gim_dkm on the GPU (csrc/dkm_api.cu through the C ABI) against golden vectors of the unmodified reference."""
import pytest
import torch

from tests.test_dkm_oracle import DKM_BIG_CASES, DKM_CASES, TOL_CERT, TOL_WARP, load_dkm_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dkm_model():
    from gim_b200 import DKMv3
    from gim_b200.dkm_params import seeded_state_dict
    m = DKMv3(None, 96, 128, upsample_preds=True)
    m.load_state_dict(seeded_state_dict(0))
    return m.eval().cuda()


def _run(m, case):
    im0, im1, h, w, up, warp, cert = load_dkm_case(case)
    m.h_resized, m.w_resized, m.upsample_res = h, w, up
    w2, c2 = m.match(im0.cuda(), im1.cuda())
    torch.cuda.synchronize()
    return w2.cpu(), c2.cpu(), warp, cert


def _cert_err(c2, cert, warp):
    """Certainty error away from the reference's hard threshold: match() zeroes the certainty where |flow| > 1
    (dkm.py:721-723), a step function of the flow - a pixel whose (clamped) warp coordinate sits within 1e-4 of +-1 may
    land on either side for a 1e-5 flow difference.  Those pixels are compared on the warp only."""
    edge = (warp.abs() >= 1 - 1e-4).any(dim=-1)
    d = (c2 - cert).abs()
    print("certainty: pixels at the |flow| = 1 step:", int(edge.sum()), "of", edge.numel(), "; max err there", d[edge].max().item() if edge.any() else 0.0)
    return d[~edge].max().item()


@pytest.mark.parametrize("case", DKM_CASES)
def test_dkm_match_vs_reference_golden(dkm_model, case):
    w2, c2, warp, cert = _run(dkm_model, case)
    assert w2.shape == warp.shape and c2.shape == cert.shape
    ew, ec = (w2 - warp).abs().max().item(), _cert_err(c2, cert, warp)
    print(case, "warp err", ew, "certainty err", ec, "launches", dkm_model.launch_count())
    assert dkm_model.launch_count() > 0
    assert ew < TOL_WARP and ec < TOL_CERT


@pytest.mark.parametrize("case", DKM_BIG_CASES)
def test_dkm_config3_geometry_vs_reference_golden(dkm_model, case):
    """BASELINE config 3 sizes (672x896, second pass 1152x1536; GP with 588 / 2352 tokens on the tensor-core Gram path and
    74 Cholesky blocks): the golden holds the reference outputs on a stride-8 grid."""
    import os
    if not os.path.isfile(os.path.join(os.path.dirname(__file__), "golden", case + ".npz")):
        pytest.skip("golden not generated")
    w2, c2, warp, cert = _run(dkm_model, case)
    w2, c2 = w2[::8, ::8], c2[::8, ::8]
    assert w2.shape == warp.shape and c2.shape == cert.shape
    ew, ec = (w2 - warp).abs().max().item(), _cert_err(c2, cert, warp)
    print(case, "warp err", ew, "certainty err", ec)
    # Tolerance at this geometry: 5e-3 / 1e-2 instead of 1e-3.  The GP at 1/16 solves a 2352 x 2352 system of condition
    # 1.9e4 in fp32, and (reference quirk, dkm.py:352-356) applies the FIRST half's inverse to the second half, which makes
    # |mu| reach 128 by cancellation.  Measured on the CPU with the reference's own matrices: its fp32 `linalg.inv` result
    # is 2.8e-3 away from the fp64 solution, an fp32 Cholesky solve 4.7e-3, the two fp32 methods 5.5e-3 apart - any two
    # fp32 algorithms differ by that much here, and the decoder carries it to ~1.6e-3 on the warp (measured).  The
    # well-conditioned cases above (up to 252 tokens) agree to 5e-6.
    assert ew < 5e-3 and ec < 1e-2


def test_dkm_sample_and_hloc_wrapper(dkm_model):
    """sample() (torch, caller's RNG) and the hloc wrapper keep the reference's output contract (dkm.py:583-620,
    hloc/matchers/dkm.py:92-154): shapes, value ranges, swapped-back keys, in-bounds keypoints."""
    from gim_b200.dkm import HlocDKM
    im0, im1, h, w, up, _, _ = load_dkm_case(DKM_CASES[1])
    dkm_model.h_resized, dkm_model.w_resized, dkm_model.upsample_res = h, w, up
    warp, cert = dkm_model.match(im0.cuda(), im1.cuda())
    torch.manual_seed(0)
    m, c = dkm_model.sample(warp, cert, 500)
    assert m.shape[1] == 4 and m.shape[0] == c.shape[0] and 0 < m.shape[0] <= 500
    assert m.abs().max().item() <= 1.0 and c.min().item() >= 0 and c.max().item() <= 1
    wrap = HlocDKM()
    wrap.net = dkm_model
    wrap.h, wrap.w = h, w
    pred = wrap({"image0": im0.cuda(), "image1": im1.cuda(), "name0": "a.jpg", "name1": "b.jpg"})
    assert set(pred) == {"keypoints0", "keypoints1", "scores", "batch_indexes"}
    k0, k1 = pred["keypoints0"], pred["keypoints1"]
    assert k0.shape == k1.shape and k0.shape[0] == pred["scores"].shape[0] > 0
    assert (k0[:, 0] <= im0.shape[3] - 1).all() and (k0[:, 1] <= im0.shape[2] - 1).all() and (k0 > 0).all()


def test_dkm_stage_taps_vs_oracle(dkm_model):
    """Pyramid levels, GP outputs and the flow after every scale against the CPU oracle (same seeded weights)."""
    from gim_b200.dkm_params import seeded_state_dict
    from oracle import dkm_oracle
    case = DKM_CASES[0]
    im0, im1, h, w, up, _, _ = load_dkm_case(case)
    taps = {}
    dkm_oracle.match(seeded_state_dict(0), im0, im1, h, w, up, taps=taps)
    names = ["dfn_flow16", "refiner_in16", "refiner_dw16", "refiner_pw16", "refiner_out16",
             "enc2", "enc4", "enc8", "enc16", "enc32", "gp32", "gp16"] + [f"flow{s}" for s in (32, 16, 8, 4, 2, 1)] + \
            [f"cert{s}" for s in (32, 16, 8, 4, 2, 1)] + [f"flow{s}u" for s in (8, 4, 2, 1)] + [f"cert{s}u" for s in (8, 4, 2, 1)]
    dkm_model.debug_taps = names
    try:
        dkm_model.h_resized, dkm_model.w_resized, dkm_model.upsample_res = h, w, up
        dkm_model.match(im0.cuda(), im1.cuda())
        torch.cuda.synchronize()
        got = {k: v.cpu() for k, v in dkm_model.last_taps.items()}
    finally:
        dkm_model.debug_taps = None
    worst = 0.0
    for k in names:
        ref = taps[k]
        ref = ref.permute(0, 2, 3, 1) if (ref.dim() == 4 and not k.startswith("cert")) else ref
        if k.startswith("cert"):
            ref = ref[:, 0]
        err = (got[k] - ref).abs().max().item()
        scale = max(1.0, ref.abs().max().item())
        print(f"{k:8s} max|ref| {ref.abs().max().item():8.3f}  err {err:.3e}")
        worst = max(worst, err / scale)
        assert err / scale < 2e-4, k


def test_kde_density_kernel_vs_reference_formula():
    """the balanced-sampling density on the device vs utils/kde.py:23-25 (cdist form)."""
    from gim_b200.dkm import RegressionMatcher
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(3001, 4, generator=g) * 2 - 1).cuda()
    ref = (-torch.cdist(x.double(), x.double()) ** 2 / (2 * 0.1 ** 2)).exp().sum(dim=-1).float()
    got = RegressionMatcher._kde(x, 0.1)
    rel = ((got - ref).abs() / ref).max().item()
    print("kde rel err", rel)
    assert rel < 1e-4
