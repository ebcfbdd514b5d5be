import sys, torch
sys.path.insert(0, '.')
from gim_b200 import DKMv3
from gim_b200.dkm_params import seeded_state_dict
from oracle import dkm_oracle
from tests.test_dkm_oracle import DKM_CASES, load_dkm_case
case = DKM_CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
im0, im1, h, w, up, warp, cert = load_dkm_case(case)
sd = seeded_state_dict(0)
taps = {}
dkm_oracle.match(sd, im0, im1, h, w, up, taps=taps)
m = DKMv3(None, h, w, upsample_preds=True); m.load_state_dict(sd); m = m.eval().cuda()
names = ["dfn_flow16", "refiner_in16", "refiner_dw16", "refiner_pw16", "refiner_out16", "enc2", "enc4", "enc8", "enc16", "enc32", "gp32", "gp16"] + [f"{k}{s}" for s in (32, 16, 8, 4, 2, 1) for k in ("flow", "cert")] + \
        [f"{k}{s}u" for s in (8, 4, 2, 1) for k in ("flow", "cert")]
m.debug_taps = names
m.upsample_res = up
w2, c2 = m.match(im0.cuda(), im1.cuda())
torch.cuda.synchronize()
for k in names:
    ref = taps[k]
    ref = ref.permute(0, 2, 3, 1) if (ref.dim() == 4 and not k.startswith("cert")) else ref
    if k.startswith("cert"): ref = ref[:, 0]
    got = m.last_taps[k].cpu()
    if k == "refiner_in16":
        for nm, a, b in (("x", 0, 512), ("x_hat", 512, 1024), ("emb", 1024, 1152), ("corr", 1152, 1377)):
            print("   refiner_in16", nm, "max|ref|", ref[..., a:b].abs().max().item(), "err", (got[..., a:b] - ref[..., a:b]).abs().max().item())
    print(f"{k:8s} shape {tuple(ref.shape)} max|ref| {ref.abs().max().item():8.3f}  err {(got-ref).abs().max().item():.3e}")
print("final warp err", (w2.cpu() - warp).abs().max().item(), "cert err", (c2.cpu() - cert).abs().max().item(), "launches", m.launch_count())
