"""Consumers either side of the matcher, host-side:

* `HlocMatchWriter` - the dense-match writer loop of hloc (`hloc/match_dense.py:240-257`): rescale keypoints to original
  image coordinates, device -> host once per pair, groups `<pair>/{keypoints0, keypoints1, scores}`.  h5py is used when it
  is importable (it is absent from this image); otherwise the same groups go into an .npz container with the same keys
  (`<pair>/keypoints0` ...), so downstream code can be pointed at either.
* ZEB result rows and file (`trainer/lightning.py:101-122, 248-275`, `tools/metrics.py:11-160`): symmetric epipolar
  errors, essential-matrix RANSAC pose (OpenCV, like the reference), rotation / translation errors and the exact text
  line format `analysis.py` / `check.py` of the reference parse.

Nothing here touches the GPU hot path; the inputs are the outputs of `LoFTR.forward` / `DKMv3.match + sample`."""
import os

import numpy as np
import torch


# ----------------------------------------------------------------------------- hloc writer (f.2)
def names_to_pair(name0, name1, separator="/"):
    """hloc/utils/parsers.py: names_to_pair."""
    return separator.join((name0.replace("/", "-"), name1.replace("/", "-")))


def scale_keypoints(kpts, scale):
    """hloc/match_dense.py: scale_keypoints - scale = original size / network size, (w, h) order."""
    if torch.is_tensor(scale):
        scale = scale.to(kpts)
    else:
        scale = kpts.new_tensor(np.asarray(scale, dtype=np.float64))
    if (scale != 1.0).any():
        kpts = kpts * scale
    return kpts


class HlocMatchWriter:
    def __init__(self, path):
        self.path = str(path)
        try:
            import h5py  # noqa: F401
            self._h5 = True
        except Exception:
            self._h5 = False
        self._npz = {}
        if not self._h5 and os.path.isfile(self.path):
            with np.load(self.path) as z:
                self._npz = {k: z[k] for k in z.files}

    def write(self, name0, name1, pred, scale0=(1.0, 1.0), scale1=(1.0, 1.0)):
        """pred: {'keypoints0', 'keypoints1', 'scores'} tensors in network-input pixels (match_dense.py:241-257)."""
        k0 = (scale_keypoints(pred["keypoints0"] + 0.5, scale0) - 0.5).cpu().numpy()
        k1 = (scale_keypoints(pred["keypoints1"] + 0.5, scale1) - 0.5).cpu().numpy()
        sc = pred["scores"].cpu().numpy()
        pair = names_to_pair(name0, name1)
        if self._h5:
            import h5py
            with h5py.File(self.path, "a") as fd:
                if pair in fd:
                    del fd[pair]
                grp = fd.create_group(pair)
                grp.create_dataset("keypoints0", data=k0)
                grp.create_dataset("keypoints1", data=k1)
                grp.create_dataset("scores", data=sc)
        else:
            self._npz[pair + "/keypoints0"], self._npz[pair + "/keypoints1"], self._npz[pair + "/scores"] = k0, k1, sc
        return pair

    def close(self):
        if not self._h5:
            with open(self.path, "wb") as f:  # exact file name (np.savez would append .npz)
                np.savez(f, **self._npz)


# ----------------------------------------------------------------------------- ZEB metrics + result file (f.3)
def cross_product_matrix(t):
    """kornia.geometry.epipolar.numeric.cross_product_matrix for [B, 3]."""
    z = torch.zeros_like(t[:, 0])
    return torch.stack([torch.stack([z, -t[:, 2], t[:, 1]], -1), torch.stack([t[:, 2], z, -t[:, 0]], -1),
                        torch.stack([-t[:, 1], t[:, 0], z], -1)], 1)


def symmetric_epipolar_distance(pts0, pts1, E, K0, K1):
    """tools/metrics.py:32-51."""
    pts0 = (pts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    pts1 = (pts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    pts0 = torch.cat([pts0, torch.ones_like(pts0[:, :1])], -1)
    pts1 = torch.cat([pts1, torch.ones_like(pts1[:, :1])], -1)
    Ep0 = pts0 @ E.T
    p1Ep0 = torch.sum(pts1 * Ep0, -1)
    Etp1 = pts1 @ E
    return p1Ep0 ** 2 * (1.0 / (Ep0[:, 0] ** 2 + Ep0[:, 1] ** 2) + 1.0 / (Etp1[:, 0] ** 2 + Etp1[:, 1] ** 2))


def relative_pose_error(T_0to1, R, t, ignore_gt_t_thr=0.0):
    """tools/metrics.py:11-29."""
    t_gt = T_0to1[:3, 3]
    n = np.linalg.norm(t) * np.linalg.norm(t_gt)
    t_err = np.rad2deg(np.arccos(np.clip(np.dot(t, t_gt) / n, -1.0, 1.0)))
    t_err = np.minimum(t_err, 180 - t_err)
    if np.linalg.norm(t_gt) < ignore_gt_t_thr:
        t_err = 0
    r = np.linalg.norm(t_gt) / np.linalg.norm(t)
    t_err2 = np.linalg.norm((t * r - t_gt))
    cos = np.clip((np.trace(np.dot(R.T, T_0to1[:3, :3])) - 1) / 2, -1.0, 1.0)
    return t_err, np.rad2deg(np.abs(np.arccos(cos))), t_err2


def estimate_pose(kpts0, kpts1, K0, K1, thresh, conf=0.99999):
    """tools/metrics.py:77-104 (OpenCV essential-matrix RANSAC + recoverPose, like the reference)."""
    import cv2
    if len(kpts0) < 5:
        return None
    kpts0 = (kpts0 - K0[[0, 1], [2, 2]][None]) / K0[[0, 1], [0, 1]][None]
    kpts1 = (kpts1 - K1[[0, 1], [2, 2]][None]) / K1[[0, 1], [0, 1]][None]
    ransac_thr = thresh / np.mean([K0[0, 0], K1[1, 1], K0[0, 0], K1[1, 1]])
    E, mask = cv2.findEssentialMat(kpts0, kpts1, np.eye(3), threshold=ransac_thr, prob=conf, method=cv2.RANSAC)
    if E is None:
        return None
    best, ret = 0, None
    for _E in np.split(E, len(E) / 3):
        n, R, t, _ = cv2.recoverPose(_E, kpts0, kpts1, np.eye(3), 1e9, mask=mask)
        if n > best:
            ret, best = (R, t[:, 0], mask.ravel() > 0), n
    return ret


def pair_metrics(mkpts0, mkpts1, K0, K1, T_0to1):
    """One pair: epipolar errors of every match + RANSAC pose errors (trainer/lightning.py:101-122 for bs = 1)."""
    K0t, K1t, Tt = (torch.as_tensor(np.asarray(a), dtype=torch.float32) for a in (K0, K1, T_0to1))
    E = cross_product_matrix(Tt[None, :3, 3])[0] @ Tt[:3, :3]
    p0, p1 = torch.as_tensor(mkpts0, dtype=torch.float32).cpu(), torch.as_tensor(mkpts1, dtype=torch.float32).cpu()
    epi = symmetric_epipolar_distance(p0, p1, E, K0t, K1t).numpy() if len(p0) else np.zeros(0, np.float32)
    ret = estimate_pose(p0.numpy(), p1.numpy(), np.asarray(K0, np.float64), np.asarray(K1, np.float64), 0.5, conf=0.99999)
    if ret is None:
        return dict(epi_errs=epi, R_err=np.inf, t_err=np.inf, t_err2=np.inf, inliers=np.zeros(0, bool))
    R, t, inl = ret
    t_err, R_err, t_err2 = relative_pose_error(np.asarray(T_0to1, np.float64), R, t)
    return dict(epi_errs=epi, R_err=R_err, t_err=t_err, t_err2=t_err2, inliers=inl)


ZEB_HEADER = "identifiers covisible0 covisible1 R_errs t_errs t_errs2 Bef.Prec Bef.Num Aft.Prec Aft.Num\n"


def zeb_result_line(identifier, covisible0, covisible1, m, eet=5e-4):
    """One line of `dump/zeb/[T] <method> <scene> <version>.txt` (trainer/lightning.py:258-270)."""
    mean = lambda x: sum(x) / max(len(x), 1)  # noqa: E731
    epi, inl = m["epi_errs"], m["inliers"]
    bef = epi < eet
    aft = epi[inl] < eet if len(inl) else np.zeros(0, bool)
    return (f"{identifier} {covisible0} {covisible1} {m['R_err']} {m['t_err']} {m['t_err2']} "
            f"{mean(bef)} {sum(bef)} {mean(aft)} {sum(aft)}\n")


def write_zeb_result_file(path, lines):
    """Rows sorted by identifier, duplicates (same identifier) removed like trainer/lightning.py:253-255."""
    uniq = {}
    for ln in lines:
        uniq.setdefault(ln.split(" ", 1)[0], ln)
    with open(path, "w") as f:
        f.write(ZEB_HEADER + "".join(uniq[k] for k in sorted(uniq)))
