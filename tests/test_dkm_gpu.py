"""gim_dkm on the GPU (csrc/dkm_api.cu through the C ABI) against golden vectors of the unmodified reference."""
import pytest
import torch

from tests.test_dkm_oracle import DKM_CASES, TOL_CERT, TOL_WARP, load_dkm_case

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


@pytest.mark.parametrize("case", DKM_CASES)
def test_dkm_match_vs_reference_golden(dkm_model, case):
    w2, c2, warp, cert = _run(dkm_model, case)
    assert w2.shape == warp.shape and c2.shape == cert.shape
    ew, ec = (w2 - warp).abs().max().item(), (c2 - cert).abs().max().item()
    print(case, "warp err", ew, "certainty err", ec, "launches", dkm_model.launch_count())
    assert dkm_model.launch_count() > 0
    assert ew < TOL_WARP and ec < TOL_CERT


def test_dkm_stage_taps_vs_oracle(dkm_model):
    """Pyramid levels, GP outputs and the flow after every scale against the CPU oracle (same seeded weights)."""
    from gim_b200.dkm_params import seeded_state_dict
    from oracle import dkm_oracle
    case = DKM_CASES[0]
    im0, im1, h, w, up, _, _ = load_dkm_case(case)
    taps = {}
    dkm_oracle.match(seeded_state_dict(0), im0, im1, h, w, up, taps=taps)
    names = ["enc2", "enc4", "enc8", "enc16", "enc32", "gp32", "gp16"] + [f"flow{s}" for s in (32, 16, 8, 4, 2, 1)] + \
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
