"""Data side of the hot path: the reference's `loader.py` surface plus a GPU-resident
batch source.

`create_npy` / `convert_npy` / `data_preprocess` / `NeRFDataset` keep the reference's names, arguments, file layout and
the tuple `__getitem__` returns (loader.py:12-133).  `GpuRayBatches` replaces `DataLoader(NeRFDataset)` for training at
GPU speed: all pixels and pose rows live in HBM and a batch is an index computation, not 4 worker processes
collating per-pixel Python objects (136 B/ray of host traffic at >1 M rays/s).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
from torch.utils.data import Dataset

NEAR_FACTOR = 2.0
FAR_FACTOR = 6.0


def create_npy(root_dir, mode):
    """loader.py:12-36: Blender transforms_<mode>.json -> <mode>.npy of [c2w(3x4) | H,W,f] rows + [near, far]."""
    from PIL import Image
    with open(root_dir + "transforms_" + mode + ".json") as fh:
        meta = json.load(fh)
    frames = meta["frames"]
    first = Image.open(root_dir + frames[0]["file_path"][2:] + ".png")
    width, height = first.size
    focal = 0.5 * width / np.tan(0.5 * meta["camera_angle_x"])
    rows = np.zeros((len(frames), 17))
    hwf = np.array([[height], [width], [focal]])
    for i, fr in enumerate(frames):
        c2w = np.array(fr["transform_matrix"])[:3, :4]
        rows[i, :15] = np.concatenate((c2w, hwf), axis=1).reshape(-1)
        rows[i, 15:] = (NEAR_FACTOR, FAR_FACTOR)
    np.save(root_dir + mode + ".npy", rows)


def convert_npy(root_dir):
    """loader.py:38-53: LLFF poses_bounds.npy -> new.npy with the rotation columns re-ordered to [c1, -c0, c2]."""
    src = np.load(root_dir + "poses_bounds.npy")
    out = np.zeros_like(src)
    for i, vec in enumerate(src):
        pose = vec[:-2].reshape(3, 5)
        rot = np.stack((pose[:, 1], -pose[:, 0], pose[:, 2]), axis=1)
        out[i, :15] = np.concatenate((rot, pose[:, 3:4], pose[:, 4:5]), axis=1).reshape(-1)
        out[i, 15:] = vec[-2:]
    np.save(root_dir + "new.npy", out)


def data_preprocess(root_dir, type, mode):
    """loader.py:55-59."""
    if type == "llff":
        convert_npy(root_dir)
    else:
        create_npy(root_dir, mode)


class NeRFDataset(Dataset):
    """loader.py:61-133: every image decoded into one (N_pix, 3) tensor; item = one pixel."""

    def __init__(self, root_dir, low_res=8, transform=None, type="sync", mode="train"):
        self.root_dir, self.low_res, self.transform, self.type = root_dir, low_res, transform, type
        trans_path = root_dir + ("new.npy" if type == "llff" else (mode + ".npy"))
        if not os.path.isfile(trans_path):
            data_preprocess(root_dir, type, mode)
        self.poses_bounds = np.load(trans_path)
        img_dir = root_dir + ("images/" if type == "llff" else (mode + "/"))
        self.file_list = [os.path.join(img_dir, f) for f in os.listdir(img_dir)]
        self.pic_num = len(self.file_list)
        self.file_list.sort(key=lambda name: int(name.split("_")[-1][:-4]))   # loader.py:111
        self.get_all_pix()

    def get_img(self, img_path):
        """loader.py:63-74: RGBA renders are composited on white."""
        from PIL import Image
        image = Image.open(img_path)
        image.load()
        if self.type == "sync":
            white = Image.new("RGB", image.size, (255, 255, 255))
            white.paste(image, mask=image.split()[3])
            image = white
        return np.array(image) / 255.0

    def get_all_pix(self):
        first = self.poses_bounds[0]
        self.height, self.width, self.focal = int(first[4]), int(first[9]), first[14]
        self.pic_size = self.height * self.width
        self.num_pix = self.pic_size * self.pic_num
        imgs = torch.zeros(self.pic_num, self.height, self.width, 3)
        for i, path in enumerate(self.file_list):
            imgs[i] = torch.tensor(self.get_img(path))
        self.all_pix = imgs.reshape(-1, 3)

    def __len__(self):
        return self.num_pix

    def __getitem__(self, idx):
        """loader.py:119-133 -> (row, column, pix_val, poses_bound, pic)."""
        pic, in_pic = divmod(idx, self.pic_size)
        row, column = divmod(in_pic, self.width)
        return row, column, self.all_pix[idx][0:3], self.poses_bounds[pic], pic


class GpuRayBatches:
    """GPU-resident replacement of DataLoader(NeRFDataset, batch_size, shuffle, drop_last=True) (nerf.py:424, 438, 442).

    Yields the collated loader tuple (row, column, pix_val, poses_bound, pic) as DEVICE tensors: int64 [B] x3, fp32
    [B,3], fp32 [B,17].  One epoch = one permutation of all pixels (shuffle) or flat order (display, nerf.py:442)."""

    def __init__(self, all_pix, poses_bounds, height, width, batch_size, shuffle=True, drop_last=True, device="cuda",
                 seed=0):
        self.dev = torch.device(device)
        self.pix = torch.as_tensor(all_pix, dtype=torch.float32).to(self.dev)
        self.poses = torch.as_tensor(np.asarray(poses_bounds), dtype=torch.float64).to(torch.float32).to(self.dev)
        self.height, self.width = int(height), int(width)
        self.pic_size = self.height * self.width
        self.n = self.pix.shape[0]
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), shuffle, drop_last
        self.gen = torch.Generator(device=self.dev if self.dev.type == "cuda" else "cpu").manual_seed(seed)

    @classmethod
    def from_dataset(cls, ds: NeRFDataset, batch_size, **kw):
        return cls(ds.all_pix, ds.poses_bounds, ds.height, ds.width, batch_size, **kw)

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def batch_of(self, idx):
        pic = torch.div(idx, self.pic_size, rounding_mode="floor")
        in_pic = idx - pic * self.pic_size
        row = torch.div(in_pic, self.width, rounding_mode="floor")
        col = in_pic - row * self.width
        return row, col, self.pix[idx], self.poses[pic], pic

    def __iter__(self):
        order = torch.randperm(self.n, generator=self.gen, device=self.dev) if self.shuffle else \
            torch.arange(self.n, device=self.dev)
        for b in range(len(self)):
            yield self.batch_of(order[b * self.batch_size:(b + 1) * self.batch_size])
