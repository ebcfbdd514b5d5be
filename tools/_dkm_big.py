import sys, torch
sys.path.insert(0, '.')
from gim_b200 import DKMv3
from gim_b200.dkm_params import seeded_state_dict
from tests.test_dkm_oracle import load_dkm_case
im0, im1, h, w, up, warp, cert = load_dkm_case("dkm_672x896_up1152x1536_s8")
taps = torch.load('tools/_dkm_big_taps.pt')
m = DKMv3(None, h, w, upsample_preds=True); m.load_state_dict(seeded_state_dict(0)); m = m.eval().cuda()
names = [k for k in taps]
m.debug_taps = names
m.upsample_res = up
w2, c2 = m.match(im0.cuda(), im1.cuda())
torch.cuda.synchronize()
for k in names:
    ref, st = taps[k]
    got = m.last_taps[k].cpu()
    if k.startswith("cert"):
        ref = ref[:, 0]; got = got[:, ::st, ::st]
    else:
        ref = ref.permute(0, 2, 3, 1); got = got[:, ::st, ::st]
    print(f"{k:10s} shape {tuple(ref.shape)} stride {st} max|ref| {ref.abs().max().item():8.3f}  err {(got-ref).abs().max().item():.3e}  nan {int(torch.isnan(got).sum())}")
d = (w2.cpu()[::8, ::8] - warp).abs()
print("final warp err", d.max().item(), "frac > 1e-3:", (d > 1e-3).float().mean().item(), "where:", (d.amax(-1) > 1e-3).nonzero()[:5].tolist())
full = w2.cpu()
print("ours  [0,1536]", full[0, 1536].tolist(), " ref[0,192]", warp[0, 192].tolist())
print("ours  [8,1544]", full[8, 1544].tolist(), " ref[1,193]", warp[1, 193].tolist())
print("ours  [0,0]", full[0, 0].tolist(), " ref[0,0]", warp[0, 0].tolist())
f1 = m.last_taps["flow1u"].cpu()
print("flow1u[1,0,0]", f1[1, 0, 0].tolist(), "flow1u[0,0,0]", f1[0, 0, 0].tolist())
