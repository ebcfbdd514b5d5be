"""world_size-2 gloo test of the pair sharding + end-of-loop gather (no GPU, no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gim_b200 import dist as gdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gdist.shard_range(n_pairs, rank, world)
    # fake per-pair results: pair k yields (k % 5) matches
    rows = []
    for k in range(lo, hi):
        m = k % 5
        data = {"m_bids": torch.zeros(m, dtype=torch.int64), "mkpts0_f": torch.full((m, 2), float(k)),
                "mkpts1_f": torch.full((m, 2), float(k) + 0.5), "mconf": torch.full((m,), 0.25)}
        rows.append(gdist.pack_matches([k], data))
    rows = torch.cat(rows, 0) if rows else torch.zeros(0, 6, dtype=torch.float64)
    counts = gdist.gather_counts(rows.shape[0])
    allrows = gdist.gather_rows(rows, dst=0)
    if rank == 0:
        torch.save({"counts": counts, "rows": allrows}, out)
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 32, 45916):
        for w in (1, 2, 4, 8):
            spans = [gdist.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather(tmp_path):
    out = str(tmp_path / "res.pt")
    n_pairs = 11
    mp.spawn(_worker, args=(2, _free_port(), n_pairs, out), nprocs=2, join=True)
    res = torch.load(out)
    expect = sum(k % 5 for k in range(n_pairs))
    assert sum(res["counts"]) == expect and len(res["counts"]) == 2
    rows = res["rows"]
    assert rows.shape == (expect, 6)
    # rank order == pair order, no duplicates, every pair present with its match count
    ids = rows[:, 0].long()
    assert torch.equal(ids, torch.sort(ids).values)
    for k in range(n_pairs):
        assert int((ids == k).sum()) == k % 5


def test_pack_matches_keeps_large_pair_ids_exact():
    big = (1 << 24) + 1  # not representable in float32
    data = {"m_bids": torch.zeros(2, dtype=torch.int64), "mkpts0_f": torch.zeros(2, 2), "mkpts1_f": torch.zeros(2, 2),
            "mconf": torch.full((2,), 0.5)}
    rows = gdist.pack_matches([big], data)
    assert rows.dtype == torch.float64 and rows[:, 0].long().tolist() == [big, big]


def test_sweep_units_and_assignment_are_deterministic_and_complete():
    """the fixed mixed-resolution list is split into work units; every unit is owned by exactly one rank
    and the greedy largest-first deal keeps the estimated load within 20 % of the mean for 2, 4 and 8 ranks."""
    from gim_b200 import sweep
    units = sweep.build_units()
    ids = [p for u in units for p in u["ids"]]
    assert ids == list(range(len(ids))) and len(ids) == 80
    for world in (1, 2, 4, 8):
        mine = sweep.assign(units, world)
        flat = sorted(i for r in mine for i in r)
        assert flat == list(range(len(units))) and mine == sweep.assign(units, world)
        load = [sum(units[i]["cost"] for i in r) for r in mine]
        assert max(load) <= 1.2 * sum(load) / world


def test_pose_pair_is_consistent_with_its_homography():
    """gim_b200.synth.pose_pair: image 1 is image 0 under the plane-induced homography of (K, R, t): points of the plane
    have zero symmetric epipolar distance under the returned pose."""
    import numpy as np
    from gim_b200 import harness, synth
    a, b, K, T = synth.pose_pair(4, 120, 160)
    assert a.shape == b.shape == (3, 120, 160)
    R, t = T[:3, :3], T[:3, 3]
    Hm = K @ (R + np.outer(t, [0, 0, 1.0]) / 4.0) @ np.linalg.inv(K)
    p0 = np.array([[20.0, 30.0], [100.0, 80.0], [140.0, 10.0], [60.0, 100.0], [80.0, 60.0], [10.0, 110.0]])
    q = (Hm @ np.concatenate([p0, np.ones((6, 1))], 1).T).T
    p1 = q[:, :2] / q[:, 2:]
    m = harness.pair_metrics(p0, p1, K, K, T)
    assert m["epi_errs"].max() < 1e-9
