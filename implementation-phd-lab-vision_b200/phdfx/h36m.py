"""Seam B for real data: the user's `Human36MPreprocessedClips` (the reference's src/dataset.py:210-437) seen as a
source of raw uint8 frames + person box, so crop / resize / normalise and the augmentation variants run on the GPU
(K1, phdfx_extract_u8[_jitter]) instead of in the DataLoader workers, and 150 KB instead of 602 KB per frame cross PCIe.

Nothing of the dataset is re-implemented here: decoding, the ground-truth caches, the crop box and the annotation
adjustments are the user's own methods and module functions, looked up at run time on the objects the entry point
imported (`dataset.Human36MPreprocessedClips` and its module).  What this adapter leaves out of the reference's
`__getitem__` is exactly the part K1 replaces: `_crop_and_resize_video_uint8`, the augmentation of the pixels and
`Normalize` (src/dataset.py:395, 405-429).
"""
from __future__ import annotations

import sys

import torch
from torch.utils.data import Dataset

_NEEDED = ("_compute_square_crop_from_2d", "_adjust_joints2d_after_crop_and_resize",
           "_adjust_camera_after_crop_and_resize")


class U8ClipDataset(Dataset):
    """item i = (frames uint8 (T,H,W,3) un-cropped, joints3d (T,17,3), joints2d (T,17,2) in crop coordinates,
    K (3,3) of the crop, box int64 (top,left,h,w)) — the same clip, index entry and annotations as `base[i]`."""

    def __init__(self, base, module=None):
        self.base = base
        self.mod = module if module is not None else sys.modules[type(base).__module__]
        missing = [n for n in _NEEDED if not hasattr(self.mod, n)]
        if missing or not hasattr(base, "_read_video_uint8_clip_fast") or not hasattr(base, "_gt_cache"):
            raise RuntimeError(f"{type(base).__name__} / {self.mod.__name__} do not look like the reference's dataset.py "
                               f"(missing {missing or 'decoder / ground-truth cache'}); use --seam a")
        self.index = base.index
        self.seq_len = base.seq_len

    def __len__(self):
        return len(self.base)

    def __getitem__(self, idx):
        b, m = self.base, self.mod
        ci = b.index[idx]
        frames = b._read_video_uint8_clip_fast(ci.video_path, ci.start, ci.end)
        n_t, H, W, C = frames.shape
        if C != 3:
            raise RuntimeError(f"{ci.video_path}: expected RGB frames, got {C} channels")
        j3_all, j2_all = b._gt_cache[ci.gt_path]
        sel = torch.arange(ci.start, ci.end, dtype=torch.long) * b.frame_skip
        if int(sel[-1]) >= j3_all.shape[0]:
            raise RuntimeError(f"joint index out of range for {ci.gt_path}: {int(sel[-1])} >= {j3_all.shape[0]}")
        j3, j2 = j3_all[sel], j2_all[sel]
        if n_t != j3.shape[0]:
            raise RuntimeError(f"{ci.video_path}: {n_t} frames but {j3.shape[0]} joint rows")
        box = m._compute_square_crop_from_2d(joints2d=j2, img_h=H, img_w=W, scale=b.crop_scale)
        j2c = m._adjust_joints2d_after_crop_and_resize(joints2d=j2, box=box, out_size=b.resize)
        K = m._adjust_camera_after_crop_and_resize(ci.cam_params, box=box, out_size=b.resize)
        return frames.contiguous(), j3, j2c, K, box
