import sys, time, torch
sys.path.insert(0, '.')
from gim_b200 import DKMv3, synth
from gim_b200.dkm_params import seeded_state_dict
m = DKMv3(None, 672, 896, upsample_preds=True); m.load_state_dict(seeded_state_dict(0)); m = m.eval().cuda()
a, b = synth.make_pairs(1, 672, 896, first=0)
a, b = a.cuda(), b.cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    torch.cuda.synchronize(); t = time.perf_counter()
    w, c = m.match(a, b)
    torch.cuda.synchronize(); print("match", i, time.perf_counter() - t, "s; launches", m.launch_count(), flush=True)
