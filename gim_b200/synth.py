"""Synthetic image-pair generator for bench.py and the parity tests.

A pair = a textured base image resized to HxW plus a seeded homography warp of it.  Uniform noise
is not a valid workload (the matcher finds nothing and the fine path is skipped), so base images
are either the small real photographs committed in tests/golden/base_images.npz or a seeded
band-limited procedural texture.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

_BASE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "tests", "golden", "base_images.npz")


def procedural_texture(h, w, seed):
    """Sum of random-orientation sinusoids + Gaussian blobs, 3 channels, values in [0, 1]."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((3, h, w), np.float32)
    for _ in range(64):
        th, fr, ph = rng.uniform(0, math.pi), rng.uniform(0.01, 0.25), rng.uniform(0, 2 * math.pi)
        amp = rng.uniform(0.2, 1.0, size=3).astype(np.float32)
        wave = np.sin((xx * math.cos(th) + yy * math.sin(th)) * fr * 2 * math.pi + ph)
        img += amp[:, None, None] * wave[None]
    for _ in range(96):
        cx, cy, s = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(3, 24)
        amp = rng.uniform(-4, 4, size=3).astype(np.float32)
        img += amp[:, None, None] * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))[None]
    img -= img.min()
    img /= max(img.max(), 1e-6)
    return torch.from_numpy(img)


def base_image(k, h, w, path=_BASE):
    """k-th base image as float [3,h,w] in [0,1] with values on the u8 grid (x/255)."""
    if os.path.isfile(path):
        z = np.load(path)
        names = sorted(z.files)
        u8 = torch.from_numpy(z[names[k % len(names)]])  # [H0,W0,3] uint8
        img = u8.permute(2, 0, 1).float()[None]
        if img.shape[-2:] != (h, w):
            img = F.interpolate(img, size=(h, w), mode="bilinear", align_corners=False, antialias=True)
        img = img[0]
        if (k // len(names)) % 2 == 1:
            img = img.flip(-1)
        return torch.round(img.clamp(0, 255)) / 255.0
    return torch.round(procedural_texture(h, w, 1234 + k) * 255.0) / 255.0


def random_homography(k):
    rng = np.random.default_rng(1000 + k)
    a, b, c, d = rng.uniform(-0.06, 0.06, size=4)
    tx, ty = rng.uniform(-12, 12, size=2)
    e, f = rng.uniform(-3e-5, 3e-5, size=2)
    return np.array([[1 + a, b, tx], [c, 1 + d, ty], [e, f, 1.0]], np.float64)


def warp_homography(img, Hm):
    """Warp [3,h,w] so that out(x') = img(H^-1 x') (bilinear, zeros outside); result on the u8 grid."""
    _, h, w = img.shape
    Hi = torch.from_numpy(np.linalg.inv(Hm)).float()
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing="ij")
    p = torch.stack([xs, ys, torch.ones_like(xs)], -1) @ Hi.T
    u, v = p[..., 0] / p[..., 2], p[..., 1] / p[..., 2]
    grid = torch.stack([(u + 0.5) / w * 2 - 1, (v + 0.5) / h * 2 - 1], -1)[None]
    out = F.grid_sample(img[None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0]
    return torch.round(out.clamp(0, 1) * 255.0) / 255.0


def make_pairs(n, h=480, w=640, first=0):
    """-> color0, color1 float32 [n,3,h,w].  Pair k uses base image (first+k) and homography (first+k)."""
    c0, c1 = [], []
    for k in range(first, first + n):
        img = base_image(k, h, w)
        c0.append(img)
        c1.append(warp_homography(img, random_homography(k)))
    return torch.stack(c0), torch.stack(c1)


def pose_pair(k, h, w):
    """A pose-consistent synthetic pair for the sweep / ZEB-format metrics: a fronto-parallel textured plane at depth 4 seen by
    two cameras with the same intrinsics and a small relative motion.  Image 1 is image 0 warped by the plane-induced
    homography H = K (R + t n^T / d) K^-1, so the ground-truth relative pose (T_0to1) and K are known exactly.
    -> img0, img1 [3,h,w] float (u8 grid), K [3,3], T_0to1 [4,4] (float64 numpy)."""
    rng = np.random.default_rng(5000 + k)
    f = 0.9 * max(h, w)
    K = np.array([[f, 0, w / 2.0], [0, f, h / 2.0], [0, 0, 1.0]])
    ax, ay, az = rng.uniform(-0.04, 0.04, size=3)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    R = Rz @ Ry @ Rx
    t = rng.uniform(-0.25, 0.25, size=3) * np.array([1.0, 1.0, 0.4])
    n, d = np.array([0.0, 0.0, 1.0]), 4.0
    Hm = K @ (R + np.outer(t, n) / d) @ np.linalg.inv(K)
    img = base_image(k, h, w)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return img, warp_homography(img, Hm / Hm[2, 2]), K, T
