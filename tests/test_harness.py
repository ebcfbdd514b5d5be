"""Host-side consumers: hloc dense-match writer layout and the ZEB result-file format / metrics."""
import os
import subprocess
import sys

import numpy as np
import torch

from gim_b200 import harness


def _pose_pair(n=400, seed=0):
    rng = np.random.default_rng(seed)
    K = np.array([[500.0, 0, 320], [0, 500.0, 240], [0, 0, 1]])
    ang = 0.1
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.3, 0.05, 0.02])
    X = np.concatenate([rng.uniform(-1, 1, (n, 2)) * 2, rng.uniform(4, 8, (n, 1))], 1)
    x0 = (K @ X.T).T
    x1 = (K @ (R @ X.T + t[:, None])).T
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t
    return x0[:, :2] / x0[:, 2:], x1[:, :2] / x1[:, 2:], K, T


def test_pair_metrics_on_exact_correspondences():
    p0, p1, K, T = _pose_pair()
    m = harness.pair_metrics(p0, p1, K, K, T)
    assert m["epi_errs"].max() < 1e-8          # exact correspondences lie on their epipolar lines
    assert m["R_err"] < 0.5 and m["t_err"] < 2.0 and m["inliers"].sum() > 300
    line = harness.zeb_result_line("scene#a#b", 0.5, 0.5, m)
    parts = line.split()
    assert len(parts) == 10 and parts[0] == "scene#a#b" and float(parts[6]) == 1.0 and int(parts[7]) == len(p0)


def test_zeb_file_is_parsed_by_the_reference_analysis_format(tmp_path):
    p0, p1, K, T = _pose_pair()
    m = harness.pair_metrics(p0, p1, K, K, T)
    lines = [harness.zeb_result_line(f"s#{i:03d}#x", 0.4, 0.6, m) for i in (2, 0, 1, 1)]
    path = tmp_path / "[T] gim_b200            GL3D test.txt"
    harness.write_zeb_result_file(path, lines)
    txt = open(path).read().splitlines()
    assert txt[0].split() == harness.ZEB_HEADER.split()
    assert [r.split()[0] for r in txt[1:]] == ["s#000#x", "s#001#x", "s#002#x"]     # sorted, de-duplicated
    # the reference's analysis.py reads `R_errs`, `t_errs` columns by header name: same columns, same order as its dumps
    ref = "/root/reference/dump/zeb"
    if os.path.isdir(ref):
        f = sorted(os.listdir(ref))[0]
        assert open(os.path.join(ref, f)).readline() == harness.ZEB_HEADER


def test_hloc_writer_layout(tmp_path):
    w = harness.HlocMatchWriter(tmp_path / "matches.h5")
    pred = {"keypoints0": torch.tensor([[10.0, 20.0], [30.0, 40.0]]), "keypoints1": torch.tensor([[1.0, 2.0], [3.0, 4.0]]),
            "scores": torch.tensor([0.9, 0.8])}
    pair = w.write("db/a.jpg", "query/b.jpg", pred, scale0=(2.0, 2.0), scale1=(1.0, 1.0))
    w.close()
    assert pair == "db-a.jpg/query-b.jpg"
    try:
        import h5py
        with h5py.File(tmp_path / "matches.h5") as fd:
            k0 = fd[pair]["keypoints0"][()]; k1 = fd[pair]["keypoints1"][()]; sc = fd[pair]["scores"][()]
    except ImportError:
        z = np.load(tmp_path / "matches.h5")
        k0, k1, sc = z[pair + "/keypoints0"], z[pair + "/keypoints1"], z[pair + "/scores"]
    assert np.allclose(k0, (np.array([[10.0, 20.0], [30.0, 40.0]]) + 0.5) * 2 - 0.5)   # match_dense.py:243
    assert np.allclose(k1, [[1.0, 2.0], [3.0, 4.0]]) and np.allclose(sc, [0.9, 0.8])
