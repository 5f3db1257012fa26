"""ORACLE (test infrastructure, never on the product path): numpy restatement of the reference's CPU front end.

Follows
  * `_crop_and_resize_video_uint8`      /root/reference/src/dataset.py:141-152
      -> torchvision `F.resize(..., antialias=False)` on a uint8 tensor
         (torchvision 0.26.0, transforms/_functional_tensor.py:441-474 resize, :516-542 _cast_squeeze_in/_out):
         cast to fp32, ATen upsample_bilinear2d (align_corners=False, scale = in/out), torch.round (half-to-even),
         cast back to uint8
      -> `.to(float32) / 255.0`
  * `self.frame_tf = T.Normalize(mean=(0.485,0.456,0.406), std=(0.229,0.224,0.225))`   dataset.py:242-245, :429
         (torchvision transforms/v2/functional/_misc.py:37-67:  (x - mean) / std in fp32)

Pinned: tests/golden/preprocess_golden.npz holds outputs of the reference functions THEMSELVES, produced in the build
container by oracle/make_golden.py (imports /root/reference/src/dataset.py through a VideoReader shim);
tests/test_oracle_preprocess.py checks this restatement against them bit-for-bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def _src_index(out_size: int, in_size: int):
    """ATen area_pixel_compute_source_index (align_corners=False, no explicit scale) + guard_index_and_lambda."""
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = scale * (dst + f32(0.5)) - f32(0.5)
    src = np.where(src < 0, f32(0), src).astype(np.float32)
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    lam1 = np.clip(src - i0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    lam0 = (f32(1) - lam1).astype(np.float32)
    i1 = np.minimum(i0 + 1, in_size - 1)
    return i0, i1, lam0, lam1


def crop_resize_u8(frames_u8: np.ndarray, box, out_size: int = 224) -> np.ndarray:
    """frames_u8: (T,H,W,3) uint8; box = (top, left, h, w) -> (T,3,out,out) uint8 (the rounded resize result).

    dataset.py:141-148.  If the crop already is out_size x out_size, F.resize returns it untouched
    (torchvision transforms/functional.py:470-471); the arithmetic below reproduces that exactly (lambda = 0).
    """
    top, left, hh, ww = (int(v) for v in box)
    crop = frames_u8[:, top:top + hh, left:left + ww, :].astype(np.float32)  # (T,h,w,3)
    y0, y1, ly0, ly1 = _src_index(out_size, hh)
    x0, x1, lx0, lx1 = _src_index(out_size, ww)
    # ATen cpu generic kernel order: inner (W) interpolation first, then H; every product and sum rounded to fp32.
    r0 = crop[:, y0]  # (T,out,w,3)
    r1 = crop[:, y1]
    lx0b = lx0[None, None, :, None]
    lx1b = lx1[None, None, :, None]
    t0 = (r0[:, :, x0] * lx0b).astype(np.float32) + (r0[:, :, x1] * lx1b).astype(np.float32)
    t1 = (r1[:, :, x0] * lx0b).astype(np.float32) + (r1[:, :, x1] * lx1b).astype(np.float32)
    ly0b = ly0[None, :, None, None]
    ly1b = ly1[None, :, None, None]
    v = (t0 * ly0b).astype(np.float32) + (t1 * ly1b).astype(np.float32)
    u8 = np.clip(np.rint(v), 0, 255).astype(np.uint8)  # torch.round = half to even, like np.rint
    return np.ascontiguousarray(u8.transpose(0, 3, 1, 2))


def crop_resize_normalize(frames_u8: np.ndarray, box, out_size: int = 224) -> np.ndarray:
    """(T,H,W,3) uint8 -> (T,3,out,out) fp32, ImageNet-normalised: dataset.py:141-152 then :429."""
    u8 = crop_resize_u8(frames_u8, box, out_size)
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - IMAGENET_MEAN[None, :, None, None]) / IMAGENET_STD[None, :, None, None]
    return x.astype(np.float32)


def hflip(video: np.ndarray) -> np.ndarray:
    """dataset.py:166 `torch.flip(video, dims=[-1])` on the resized clip."""
    return np.ascontiguousarray(video[..., ::-1])


def to_nhwc4p_bf16_bits(x_nchw_f32: np.ndarray) -> np.ndarray:
    """fp32 (N,3,224,224) -> uint16 bf16 bit patterns in the library's NHWC4p layout (N,224,232,4), RNE rounding."""
    n = x_nchw_f32.shape[0]
    bits = x_nchw_f32.astype(np.float32).view(np.uint32)
    rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16).astype(np.uint16)
    out = np.zeros((n, 224, 232, 4), dtype=np.uint16)
    out[:, :, 4:228, 0:3] = rounded.transpose(0, 2, 3, 1)
    return out
