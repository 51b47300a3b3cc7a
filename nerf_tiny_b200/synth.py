"""Synthetic poses / pixels shaped like the reference loader's batches.

There are no datasets in the build or GPU containers, so benchmarks and tests
use an analytic scene: cameras on a radius-4 sphere looking
at the origin, a unit sphere shaded by its normal on a white background.  A
batch has the layout `NeRFDataset.__getitem__` collates to (loader.py:119-133):
row, column int64 [N]; pix_val fp32 [N,3]; poses_bound fp64 [N,17]; pic int64 [N].
Pure numpy/torch host code, no CUDA.
"""
from __future__ import annotations

import math

import numpy as np
import torch

CAMERA_ANGLE_X = 0.6911112070083618   # nerf_synthetic/lego transforms_*.json


def focal_of(width: float, angle: float = CAMERA_ANGLE_X) -> float:
    """loader.py:23."""
    return 0.5 * width / math.tan(0.5 * angle)


def sphere_pose(theta: float, phi: float, radius: float = 4.0) -> np.ndarray:
    """3x4 camera-to-world, columns [right, up, back | position] (camera looks down -z)."""
    pos = np.array([radius * math.cos(phi) * math.cos(theta), radius * math.cos(phi) * math.sin(theta),
                    radius * math.sin(phi)])
    back = pos / np.linalg.norm(pos)
    right = np.cross(np.array([0.0, 0.0, 1.0]), back)
    right /= np.linalg.norm(right)
    up = np.cross(back, right)
    return np.stack((right, up, back, pos), axis=1)


def pose_rows(n_pose: int, height: int, width: int, focal: float, near=2.0, far=6.0, seed: int = 0,
              llff_bounds: bool = False) -> np.ndarray:
    """[n_pose,17] float64 rows as `create_npy` writes them (loader.py:33): c2w(3x4)|H,W,f + near, far.
    With llff_bounds the near/far differ per image (LLFF-like, no NDC; loader.py:38-53)."""
    rng = np.random.RandomState(seed)
    rows = np.zeros((n_pose, 17))
    for i in range(n_pose):
        theta = 2 * math.pi * i / n_pose
        phi = 0.25 + 0.35 * math.sin(1.7 * i)
        c2w = sphere_pose(theta, phi)
        nr, fr = (rng.uniform(1.0, 1.6), rng.uniform(8.0, 16.0)) if llff_bounds else (near, far)
        rows[i] = np.concatenate((np.concatenate((c2w, np.array([[height], [width], [focal]])), axis=1).flatten(),
                                  np.array([nr, fr])))
    return rows


def k_inv_of(height: float, width: float, focal: float) -> torch.Tensor:
    """nerf.py:433 (transposed inverse intrinsics)."""
    return torch.tensor([[1.0, 0.0, -0.5 * width], [0.0, -1.0, 0.5 * height], [0.0, 0.0, -focal]]
                        ).to(torch.float).transpose(0, 1).contiguous()


def shade(rows17: np.ndarray, pic: np.ndarray, row: np.ndarray, col: np.ndarray) -> np.ndarray:
    """Target colours of the analytic scene for pixels (pic,row,col) under the reference's
    x:=row, y:=column convention (nerf.py:343-344, 433)."""
    pose = rows17[pic, :15].reshape(-1, 3, 5)
    h, w, f = pose[:, 0, 4], pose[:, 1, 4], pose[:, 2, 4]
    v = np.stack((row - 0.5 * w, -col + 0.5 * h, -f), axis=1)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    d = np.einsum("nrc,nc->nr", pose[:, :, :3], v)
    o = pose[:, :, 3]
    b = np.sum(o * d, axis=1)
    c = np.sum(o * o, axis=1) - 1.0
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0.0))
    nrm = o + t[:, None] * d
    colour = np.where(hit[:, None], 0.5 * (nrm + 1.0), 1.0)
    return colour.astype(np.float32)


def view_batch(rows17: np.ndarray, pic: int, height: int, width: int):
    """All pixels of one view in the loader's flat order (loader.py:123-130)."""
    idx = np.arange(height * width)
    row = idx // width
    col = idx % width
    return _pack(rows17, np.full_like(idx, pic), row, col)


def random_batch(rows17: np.ndarray, n_rays: int, height: int, width: int, gen: torch.Generator):
    """`n_rays` pixels uniform over (pose,row,col) — what DataLoader(shuffle=True) yields (nerf.py:424).
    NB the reference uses `row` as the horizontal coordinate against W, so row must stay < min(H,W)
    only in the sense of image bounds; we sample row<H, col<W like the loader."""
    n_pose = rows17.shape[0]
    pic = torch.randint(0, n_pose, (n_rays,), generator=gen).numpy()
    row = torch.randint(0, height, (n_rays,), generator=gen).numpy()
    col = torch.randint(0, width, (n_rays,), generator=gen).numpy()
    return _pack(rows17, pic, row, col)


def _pack(rows17, pic, row, col):
    pix = torch.from_numpy(shade(rows17, pic, row.astype(np.float64), col.astype(np.float64)))
    return (torch.from_numpy(row.astype(np.int64)), torch.from_numpy(col.astype(np.int64)), pix,
            torch.from_numpy(rows17[pic]), torch.from_numpy(pic.astype(np.int64)))
