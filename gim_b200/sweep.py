"""ZEB-style sweep over a FIXED mixed-resolution pair list, sharded per work unit across the GPUs of one box
(BASELINE config 5, SURVEY.md section 8 d / e):

    python -m gim_b200.sweep --out gpurun_out/sweep                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m gim_b200.sweep --out ...

The ZEB data is absent, so the list is synthetic but keeps the benchmark's geometry
(TEST_GIM_LOFTR.sh:3-14, datasets/*/__init__.py): 480x640 batch-1 and batch-8 sets, the KITTI set (376x1240 content
zero-padded to 1240x1240 with padding masks, batch 8) and the 1600-wide ETH3D set (padded to 1600x1600, batch 1).
Every pair is pose-consistent (gim_b200.synth.pose_pair), so the reference's per-pair metrics are defined.

Work units (one forward each) are dealt to the ranks largest-first (greedy on the estimated cost), each rank runs its
units through the uint8 host entry (`LoFTR.forward_u8`: H2D of bytes, GPU pre-processing, forward, D2H), computes the
reference's pair metrics on the host, and at the end the ranks exchange ONE int64 match count (NCCL all_gather) plus a
variable-length gather of the result rows to rank 0, which writes the ZEB result file (`dump/zeb` line format) and a
JSON summary with per-rank pairs / time and the imbalance.  Scaling is STRONG: the list is the same for every N."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import dist as gdist
from . import harness, synth

# name, content (h, w), padded (H, W), batch, pairs
SETS = [
    ("GL3D", (480, 640), (480, 640), 1, 24),
    ("GTASfM", (480, 640), (480, 640), 8, 32),
    ("KITTI", (376, 1240), (1240, 1240), 2, 8),   # the reference runs this set with batch 8; units of 2 keep 8 GPUs balanced
    ("RobotcarNight", (768, 1024), (768, 1024), 1, 12),
    ("ETH3DO", (1064, 1600), (1600, 1600), 1, 4),
]


def build_units(scale=1.0):
    """-> list of work units {set, ids, content, padded, cost}; `scale` shrinks the list for smoke tests."""
    units, pid = [], 0
    for name, content, padded, batch, count in SETS:
        count = max(batch, int(round(count * scale)) // batch * batch)
        for s in range(0, count, batch):
            ids = list(range(pid + s, pid + s + batch))
            cost = batch * padded[0] * padded[1] * (1.0 + padded[0] * padded[1] / (64 * 4800 * 40.0))  # convs + L*S term
            units.append(dict(set=name, ids=ids, content=content, padded=padded, cost=cost))
        pid += count
    return units


def assign(units, world):
    """Greedy largest-first assignment -> per-rank lists of unit indices (deterministic)."""
    order = sorted(range(len(units)), key=lambda i: (-units[i]["cost"], i))
    load, mine = [0.0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (load[j], j))
        mine[r].append(i)
        load[r] += units[i]["cost"]
    return mine


def make_unit_inputs(u):
    h, w = u["content"]
    H, W = u["padded"]
    n = len(u["ids"])
    u0 = torch.zeros(n, h, w, 3, dtype=torch.uint8)
    u1 = torch.zeros(n, h, w, 3, dtype=torch.uint8)
    Ks, Ts = [], []
    for k, pid in enumerate(u["ids"]):
        a, b, K, T = synth.pose_pair(pid, h, w)
        u0[k] = torch.round(a * 255).to(torch.uint8).permute(1, 2, 0)
        u1[k] = torch.round(b * 255).to(torch.uint8).permute(1, 2, 0)
        Ks.append(K); Ts.append(T)
    data = {"color0_u8": u0.pin_memory() if torch.cuda.is_available() else u0, "color1_u8": u1.pin_memory() if torch.cuda.is_available() else u1}
    if (H, W) != (h, w):
        data["pad0"], data["pad1"] = (H, W), (H, W)
    return data, Ks, Ts


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the pair list (smoke tests)")
    ap.add_argument("--repeat", type=int, default=1, help="timed passes over the rank's units (after one warm-up pass)")
    args = ap.parse_args(argv)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("gim_b200.sweep needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from . import LoFTR, get_default_config, load_default_weights
    model = LoFTR(get_default_config())
    model.load_state_dict(load_default_weights())
    model = model.eval().to(dev)

    units = build_units(args.scale)
    mine = assign(units, world)[rank]
    inputs = {i: make_unit_inputs(units[i]) for i in mine}   # host-side generation is NOT timed (the ZEB loader's job)
    for i in mine:                                          # warm-up pass: workspaces, PE tables, lazy init
        model.forward_u8(dict(inputs[i][0]))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # ---- timed: the matcher end to end (uint8 host images in, matches on the host out) over this rank's units
    t0 = time.perf_counter()
    outs = {}
    order = [i for _ in range(args.repeat) for i in mine]
    staged = model.stage_u8(inputs[order[0]][0]) if order else None
    for k, i in enumerate(order):
        # the upload + GPU pre-processing of the next unit runs on the copy stream while this unit's forward computes
        nxt = model.stage_u8(inputs[order[k + 1]][0]) if k + 1 < len(order) else None
        d = dict(inputs[i][0], staged=staged)
        model.forward_u8(d)
        d.pop("staged")
        outs[i] = d
        staged = nxt
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.repeat
    # ---- not timed with the matcher: the reference's host-side pair metrics (OpenCV RANSAC, ~0.2 s per pair)
    t1 = time.perf_counter()
    rows, lines, M_rank = [], [], 0
    for i in mine:
        d, (_, Ks, Ts) = outs[i], inputs[i]
        M_rank += int(d["b_ids"].numel())
        rows.append(gdist.pack_matches(units[i]["ids"], d))
        for b, pid in enumerate(units[i]["ids"]):
            sel = d["m_bids"] == b
            m = harness.pair_metrics(d["mkpts0_f"][sel], d["mkpts1_f"][sel], Ks[b], Ks[b], Ts[b])
            lines.append(harness.zeb_result_line(f"{units[i]['set']}#{pid:08d}#{pid:08d}", 1.0, 1.0, m))
    metrics_s = time.perf_counter() - t1
    # ---- the one collective: match counts (NCCL all_gather), then the variable-length row gather to rank 0
    counts = gdist.gather_counts(M_rank, device=dev)
    allrows = gdist.gather_rows(torch.cat(rows).to(dev) if rows else torch.zeros(0, 6, dtype=torch.float64, device=dev))
    stats = torch.tensor([wall, float(sum(len(units[i]["ids"]) for i in mine)), metrics_s], dtype=torch.float64, device=dev)
    if world > 1:
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
        gl = [None] * world
        dist.all_gather_object(gl, lines)  # a few KB of text lines; the reference pickles whole metric dicts here (tools/comm.py:141)
        lines = [ln for part in gl for ln in part]
    else:
        allstats = [stats]
    if rank == 0:
        os.makedirs(args.out, exist_ok=True)
        harness.write_zeb_result_file(os.path.join(args.out, f"[T] gim_b200_loftr        SYNTH-ZEB n{world}.txt"), lines)
        walls = [float(s[0]) for s in allstats]
        pairs = [int(s[1]) for s in allstats]
        total_pairs = sum(pairs)
        R = np.array([float(ln.split()[3]) for ln in lines])
        summary = {
            "workload": "synthetic ZEB-geometry pair list (fixed): " + ", ".join(f"{n} {c[0]}x{c[1]}->{p[0]}x{p[1]} b{b} x{k}" for n, c, p, b, k in SETS),
            "n_gpus": world, "scaling": "strong", "pairs": total_pairs, "units": len(units), "matches": int(sum(counts)),
            "time_s_max_over_ranks": max(walls), "pairs_per_s": total_pairs / max(walls),
            "per_rank_pairs": pairs, "per_rank_time_s": walls, "imbalance": max(walls) / (sum(walls) / len(walls)),
            "rows_gathered": int(allrows.shape[0]) if allrows is not None else None,
            "host_metrics_s_max_over_ranks": max(float(s[2]) for s in allstats),
            "pose_R_err_deg_median": float(np.median(R[np.isfinite(R)])) if np.isfinite(R).any() else None,
            "pose_ok_frac_5deg": float((R < 5).mean()),
            "timed": "host wall clock per rank around stage_u8 + forward_u8 of its units (H2D of uint8 images + GPU pre-processing + forward + D2H; the upload of unit i+1 overlaps the forward of unit i), max over ranks; the reference's host-side pair metrics (OpenCV RANSAC) are reported separately",
        }
        with open(os.path.join(args.out, f"sweep_n{world}.json"), "w") as f:
            json.dump(summary, f, indent=1)
        print(json.dumps(summary), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
