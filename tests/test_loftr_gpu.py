"""GPU parity: the CUDA path (through the C ABI, via the drop-in module) against
 (a) golden vectors produced by the unmodified reference, and (b) the CPU oracle run live."""
import pytest
import torch

from tests.goldens import BIG_CASES, CASES, assert_matches_equal, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    from gim_b200 import LoFTR, get_default_config, load_default_weights
    m = LoFTR(get_default_config())
    m.load_state_dict(load_default_weights())
    return m.eval().cuda()


def to_cuda(data):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}


@pytest.mark.parametrize("case", CASES)
def test_golden_parity(model, case):
    data, gold = load_case(case)
    d = to_cuda(data)
    model(d)
    errs = assert_matches_equal(d, gold, what=case + ": ")
    print(case, "M =", d["b_ids"].numel(), errs)
    assert d["mkpts0_f"].dtype == torch.float32 and d["b_ids"].dtype == torch.int64
    assert d["hw0_c"] == torch.Size((data["color0"].shape[2] // 8, data["color0"].shape[3] // 8))


@pytest.mark.parametrize("case", BIG_CASES)
def test_golden_parity_big_configs(model, case):
    """BASELINE.json configs beyond 480x640, against goldens of the unmodified reference: config 1 (demo pair at
    1000x1000: L = S = 15625 = 122 full m-tiles + a 9-row tail), the ZEB KITTI geometry (1240x1240 zero-padded, masks
    + scales, batch 2, L = 24025) and the ZEB ETH3D geometry (1600 wide, L = 26600)."""
    data, gold = load_case(case)
    d = to_cuda(data)
    model(d)
    errs = assert_matches_equal(d, gold, what=case + ": ")
    print(case, "M =", d["b_ids"].numel(), errs)


def test_headline_batch_32_vs_reference_goldens(model):
    """The bench workload itself (32 pairs @ 480x640, synth.make_pairs(32, first=0)): pairs 0, 1, 13 and 31 of the batch
    against goldens of the unmodified reference (the reference treats pairs independently)."""
    from gim_b200 import synth
    c0, c1 = synth.make_pairs(32, 480, 640, first=0)
    stored, _ = load_case("synth_b2_480x640")  # pairs 0, 1 with their pixels stored: the generator must reproduce them
    assert torch.equal(c0[:2], stored["color0"]) and torch.equal(c1[:2], stored["color1"]), "synth is not reproducible here"
    d = dict(color0=c0.cuda(), color1=c1.cuda(), image0=c0, image1=c1)
    model(d)
    assert d["b_ids"].numel() > 100000
    keys = ("b_ids", "i_ids", "j_ids", "m_bids", "mconf", "mkpts0_c", "mkpts1_c", "mkpts0_f", "mkpts1_f", "expec_f")
    for case, pairs in (("synth_b2_480x640", (0, 1)), ("synth_p13_p31", (13, 31))):
        _, gold = load_case(case)
        for local, p in enumerate(pairs):
            sel = d["b_ids"] == p
            gsel = gold["b_ids"] == local
            out = {k: d[k][sel].cpu() for k in keys}
            out["b_ids"] = out["b_ids"] - p + local
            out["m_bids"] = out["m_bids"] - p + local
            errs = assert_matches_equal(out, {k: gold[k][gsel] for k in keys}, what=f"batch-32 pair {p}: ")
            print("pair", p, "M =", int(sel.sum()), errs)


def test_tc_conf_matrix_vs_oracle_240x320(model):
    """The confidence matrix written by the tcgen05 conf sweep (the product path's own sweeps, debug tap) against the
    oracle at a size with several tiles per dimension (L = S = 1200) and a batch of 2."""
    from gim_b200 import load_default_weights
    from oracle import loftr_oracle
    data, gold = load_case("small_b2_240x320")
    ref = loftr_oracle.loftr_forward(load_default_weights(), data, return_intermediates=True)
    d = to_cuda(data)
    d["return_conf_matrix"] = True
    model(d)
    assert_matches_equal(d, gold, what="with conf tap: ")  # same ids as without the tap: one numerical path
    err = (d["conf_matrix"].cpu() - ref["_inter"]["conf_matrix"]).abs().max().item()
    print("tc conf_matrix err", err)
    # entries are exp() of sims of magnitude ~10 (rel. 1e-6 per operand): measured 1.3e-5 on entries near 1; mconf's bar is 1e-3
    assert err < 5e-5


def test_stage_taps_tiny(model):
    """Stage-level parity on the tiny case: backbone maps, transformer outputs, fine windows."""
    data, gold = load_case("tiny_64x96")
    d = to_cuda(data)
    model(d, taps=["feat_c_backbone0", "feat_c_backbone1", "feat_f0", "feat_f1", "feat_c0", "feat_c1",
                   "fine_win0", "fine_win1", "conf_matrix"])
    t = {k: v.cpu() for k, v in d["_taps"].items()}
    fc = torch.cat([t["feat_c_backbone0"], t["feat_c_backbone1"]], 0).permute(0, 3, 1, 2)
    ff = torch.cat([t["feat_f0"], t["feat_f1"]], 0).permute(0, 3, 1, 2)
    e_c = (fc - gold["inter_feat_c_backbone"]).abs().max().item()
    e_f = (ff - gold["inter_feat_f"]).abs().max().item()
    e_t0 = (t["feat_c0"] - gold["inter_feat_c0"]).abs().max().item()
    e_t1 = (t["feat_c1"] - gold["inter_feat_c1"]).abs().max().item()
    print("backbone c/f err", e_c, e_f, "transformer err", e_t0, e_t1)
    assert e_c < 1e-4 and e_f < 1e-4
    assert e_t0 < 2e-4 and e_t1 < 2e-4
    if gold["b_ids"].numel():
        e_w0 = (t["fine_win0"] - gold["inter_fine_win0"]).abs().max().item()
        e_w1 = (t["fine_win1"] - gold["inter_fine_win1"]).abs().max().item()
        print("fine windows err", e_w0, e_w1)
        assert e_w0 < 2e-4 and e_w1 < 2e-4


def test_conf_matrix_vs_oracle(model):
    from gim_b200 import load_default_weights
    from oracle import loftr_oracle
    data, _ = load_case("tiny_64x96")
    ref = loftr_oracle.loftr_forward(load_default_weights(), data, return_intermediates=True)
    d = to_cuda(data)
    d["return_conf_matrix"] = True
    model(d)
    err = (d["conf_matrix"].cpu() - ref["_inter"]["conf_matrix"]).abs().max().item()
    print("conf_matrix err", err)
    # fp32 noise floor of a dual softmax over logits of magnitude ~60: ulp(60)/2 = 1.9e-6 absolute on every logit,
    # i.e. ~2e-6 relative on every exponential; the reference's own conf moves by 1e-5 with the accumulation order
    assert err < 3e-5


def test_host_entry_matches_device_entry(model):
    data, gold = load_case("small_b2_240x320")
    d = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
    model(d)  # CPU tensors -> gimb_loftr_forward_host
    assert d["mkpts1_f"].device.type == "cpu"
    assert_matches_equal(d, gold, what="host entry: ")
    assert model.last_h2d_bytes == 2 * data["color0"].numel() * 4
    assert model.last_d2h_bytes > 0


def test_live_oracle_random_pair(model):
    """A pair that has no committed golden: oracle run live on the GPU box's CPU."""
    from gim_b200 import load_default_weights, synth
    from oracle import loftr_oracle
    c0, c1 = synth.make_pairs(1, 160, 224, first=11)
    ref = loftr_oracle.loftr_forward(load_default_weights(), dict(color0=c0, color1=c1))
    d = dict(color0=c0.cuda(), color1=c1.cuda(), image0=c0, image1=c1)
    model(d)
    assert_matches_equal(d, {k: ref[k] for k in ("b_ids", "i_ids", "j_ids", "m_bids", "mconf", "mkpts0_c", "mkpts1_c",
                                                  "mkpts0_f", "mkpts1_f", "expec_f")}, what="live oracle: ")


def test_errors_are_loud(model):
    c = torch.zeros(1, 3, 100, 128, device="cuda")
    with pytest.raises(RuntimeError):
        model(dict(color0=c, color1=c, image0=c, image1=c))  # 100 is not a multiple of 8


def test_batch_consistency_and_fine_chunking(model):
    """Pairs are independent: a batch of 6 pairs (M > 16384, so the fine stage runs in more than one chunk) must give
    exactly the rows that three batches of 2 give (size-independent property at the bench resolution)."""
    from gim_b200 import synth
    c0, c1 = synth.make_pairs(6, 480, 640, first=0)
    big = dict(color0=c0.cuda(), color1=c1.cuda(), image0=c0, image1=c1)
    model(big)
    assert big["b_ids"].numel() > 16384, "workload too easy to exercise the chunk loop"
    off = 0
    for k in range(0, 6, 2):
        d = dict(color0=c0[k:k + 2].cuda(), color1=c1[k:k + 2].cuda(), image0=c0[k:k + 2], image1=c1[k:k + 2])
        model(d)
        m = d["b_ids"].numel()
        sl = slice(off, off + m)
        assert torch.equal(big["b_ids"][sl], d["b_ids"] + k)
        for key in ("i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f", "expec_f"):
            assert torch.equal(big[key][sl], d[key]), key
        off += m
    assert off == big["b_ids"].numel()
    # ordered by (b, i) like torch.where
    key = big["b_ids"] * 4800 + big["i_ids"]
    assert bool((key[1:] > key[:-1]).all())


def test_u8_entry_matches_float_entry_and_masks(model):
    """uint8 HWC host images -> device /255, CHW, zero padding and 1/8 padding masks; the same matches as
    the float entry fed with the host-side pre-processing of the reference loader (datasets/utils.py:112-124)."""
    import numpy as np
    z = np.load("tests/golden/small_b2_240x320.npz")
    u0 = torch.from_numpy(z["color0_u8"]).permute(0, 2, 3, 1).contiguous()  # [N, h, w, 3] uint8
    u1 = torch.from_numpy(z["color1_u8"]).permute(0, 2, 3, 1).contiguous()
    n, h, w, _ = u0.shape
    # (a) no padding: identical to the golden
    d = {"color0_u8": u0, "color1_u8": u1}
    model.forward_u8(d)
    _, gold = load_case("small_b2_240x320")
    assert_matches_equal(d, gold, what="u8 entry vs golden: ")
    assert model.last_h2d_bytes == 2 * n * h * w * 3
    # (b) padded to a square with masks: equals the float entry on host-padded inputs + host-made masks
    P = 320
    d2 = {"color0_u8": u0, "color1_u8": u1, "pad0": (P, P), "pad1": (P, P)}
    model.forward_u8(d2)
    c0 = torch.zeros(n, 3, P, P); c1 = torch.zeros(n, 3, P, P)
    c0[:, :, :h, :w] = u0.permute(0, 3, 1, 2).float() / 255
    c1[:, :, :h, :w] = u1.permute(0, 3, 1, 2).float() / 255
    m = torch.zeros(n, P // 8, P // 8, dtype=torch.bool); m[:, : h // 8, : w // 8] = True
    ref = to_cuda({"color0": c0, "color1": c1, "image0": c0, "image1": c1, "mask0": m, "mask1": m.clone()})
    model(ref)
    for k in ("b_ids", "i_ids", "j_ids"):
        assert torch.equal(d2[k], ref[k].cpu()), k
    assert (d2["mkpts1_f"] - ref["mkpts1_f"].cpu()).abs().max().item() < 1e-4
    assert d2["b_ids"].numel() > 100


def test_staged_u8_uploads_overlap_and_match(model):
    """stage_u8 (upload + pre-processing on the model's copy stream, no wait) then forward_u8(staged=...): two batches staged
    ahead of their forwards give the same results as the one-call entry; a staged batch cannot be consumed twice."""
    import numpy as np
    z = np.load("tests/golden/small_b2_240x320.npz")
    u0 = torch.from_numpy(z["color0_u8"]).permute(0, 2, 3, 1).contiguous().pin_memory()
    u1 = torch.from_numpy(z["color1_u8"]).permute(0, 2, 3, 1).contiguous().pin_memory()
    _, gold = load_case("small_b2_240x320")
    s_a = model.stage_u8({"color0_u8": u0, "color1_u8": u1})
    s_b = model.stage_u8({"color0_u8": u1, "color1_u8": u0, "pad0": (320, 320), "pad1": (320, 320)})   # swapped + padded
    d_a = {"staged": s_a}
    model.forward_u8(d_a)
    assert_matches_equal(d_a, gold, what="staged u8 vs golden: ")
    d_b = {"staged": s_b}
    model.forward_u8(d_b)
    ref = {"color0_u8": u1, "color1_u8": u0, "pad0": (320, 320), "pad1": (320, 320)}
    model.forward_u8(ref)
    for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts0_f", "mkpts1_f"):
        assert torch.equal(d_b[k], ref[k]), k
    assert d_b["b_ids"].numel() > 100 and model.last_h2d_bytes == 2 * u0.numel()
    with pytest.raises(RuntimeError, match="already been consumed"):
        model.forward_u8({"staged": s_a})


def test_corr_range_flag_falls_back_to_exact_sweeps():
    """csrc/corr_sweep.cu: a softmax sum outside the safe range of the fixed exponent reference raises the device flag and
    the forward repeats the coarse matching with the exact (online-max) sweeps.  Forced here by biasing the reference by
    +200 (every exponential flushes to zero); the results must still equal the golden and the fallback must be counted.
    A subprocess, because the bias is read once per process."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, sys; sys.path.insert(0, '.')\n"
        "from gim_b200 import LoFTR, get_default_config, load_default_weights\n"
        "from tests.goldens import load_case, assert_matches_equal\n"
        "m = LoFTR(get_default_config()); m.load_state_dict(load_default_weights()); m = m.eval().cuda()\n"
        "data, gold = load_case('small_b2_240x320')\n"
        "d = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}\n"
        "m(d); torch.cuda.synchronize()\n"
        "assert_matches_equal(d, gold)\n"
        "print('FALLBACKS', m.corr_fallbacks())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for bias, expect in (("200", True), ("0", False)):
        env = dict(os.environ, GIMB_CORR_GBIAS=bias)
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        n = int(r.stdout.strip().split("FALLBACKS")[-1])
        assert (n > 0) == expect, (bias, n)
