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


def _fma(a, b, c):
    """fp32 fused multiply-add: the product of two fp32 values is exact in fp64, one rounding to fp32 follows."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(np.float32)


def _mul(a, b):
    return (np.asarray(a, np.float32) * np.asarray(b, np.float32)).astype(np.float32)


def _src_index(out_size: int, in_size: int):
    """ATen area_pixel_compute_source_index (align_corners=False, no explicit scale) + guard_index_and_lambda.
    The shipped ATen binary contracts scale*(dst+0.5)-0.5 into one FMA (established empirically, see header)."""
    f32 = np.float32
    scale = f32(in_size) / f32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = _fma(np.full_like(dst, scale), dst + f32(0.5), np.full_like(dst, f32(-0.5)))
    src = np.where(src < 0, f32(0), src).astype(np.float32)
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    lam1 = np.clip(src - i0.astype(np.float32), f32(0), f32(1)).astype(np.float32)
    lam0 = (f32(1) - lam1).astype(np.float32)
    i1 = np.minimum(i0 + 1, in_size - 1)
    return i0, i1, lam0, lam1


def crop_resize_u8(frames_u8: np.ndarray, box, out_size: int = 224, aten_path: str = "worker") -> np.ndarray:
    """frames_u8: (T,H,W,3) uint8; box = (top, left, h, w) -> (T,3,out,out) uint8 (the rounded resize result).

    dataset.py:141-148.  If the crop already is out_size x out_size, F.resize returns it untouched
    (torchvision transforms/functional.py:470-471); the arithmetic below reproduces that exactly (lambda = 0).

    torch 2.11's CPU upsample_bilinear2d picks one of two kernels for 3-channel fp32 input
    (ATen UpSampleKernel.cpp, upsample_bilinear2d_kernel_impl_float), which differ in the last ulp:
      aten_path="worker"  : get_num_threads() == 1 — what the reference's DataLoader workers run
                            (src/preprocess_resnet_features.py:195-204, num_workers=8): the channels-last kernel,
                            v = fma(w11,v11, fma(w10,v10, fma(w00,v00, w01*v01))),  wij = lam_h_i * lam_w_j
      aten_path="generic" : multi-threaded caller (num_workers=0): the separable generic kernel,
                            t_i = fma(v_i0, lw0, v_i1*lw1);  v = fma(t_0, lh0, t_1*lh1)
    Both orders were established by bit-exact comparison with the installed ATen (oracle/make_golden.py pins them).
    The CUDA kernel implements "worker".
    """
    top, left, hh, ww = (int(v) for v in box)
    crop = frames_u8[:, top:top + hh, left:left + ww, :].astype(np.float32)  # (T,h,w,3)
    y0, y1, ly0, ly1 = _src_index(out_size, hh)
    x0, x1, lx0, lx1 = _src_index(out_size, ww)
    r0 = crop[:, y0]  # (T,out,w,3)
    r1 = crop[:, y1]
    v00, v01, v10, v11 = r0[:, :, x0], r0[:, :, x1], r1[:, :, x0], r1[:, :, x1]
    a = lx0[None, None, :, None]
    b = lx1[None, None, :, None]
    c = ly0[None, :, None, None]
    d = ly1[None, :, None, None]
    if aten_path == "worker":
        w00, w01, w10, w11 = _mul(c, a), _mul(c, b), _mul(d, a), _mul(d, b)
        v = _fma(w11, v11, _fma(w10, v10, _fma(w00, v00, _mul(w01, v01))))
    elif aten_path == "generic":
        t0 = _fma(v00, a, _mul(v01, b))
        t1 = _fma(v10, a, _mul(v11, b))
        v = _fma(t0, c, _mul(t1, d))
    else:
        raise ValueError(aten_path)
    u8 = np.clip(np.rint(v), 0, 255).astype(np.uint8)  # torch.round = half to even, like np.rint
    return np.ascontiguousarray(u8.transpose(0, 3, 1, 2))


def crop_resize_normalize(frames_u8: np.ndarray, box, out_size: int = 224, aten_path: str = "worker") -> np.ndarray:
    """(T,H,W,3) uint8 -> (T,3,out,out) fp32, ImageNet-normalised: dataset.py:141-152 then :429."""
    u8 = crop_resize_u8(frames_u8, box, out_size, aten_path)
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - IMAGENET_MEAN[None, :, None, None]) / IMAGENET_STD[None, :, None, None]
    return x.astype(np.float32)


def hflip(video: np.ndarray) -> np.ndarray:
    """dataset.py:166 `torch.flip(video, dims=[-1])` on the resized clip."""
    return np.ascontiguousarray(video[..., ::-1])


# ---------------------------------------------------------------------------------------------------- colour jitter
# /root/reference/src/dataset.py:188-198 `_aug_color_jitter`: T.ColorJitter(brightness=0.3, contrast=0.3,
# saturation=0.2, hue=0.05) applied ONCE to the whole (T,3,224,224) clip in [0,1] (one parameter draw per clip),
# before Normalize (:413-429).  torchvision 0.26.0 transforms/v2/_color.py:146-173 (make_params: fn_idx =
# randperm(4), factors uniform; transform: the four ops in fn_idx order: 0 brightness, 1 contrast, 2 saturation,
# 3 hue) and transforms/v2/functional/_color.py: _rgb_to_grayscale_image :31-48, _blend :92-97,
# adjust_brightness_image :112-123, adjust_saturation_image :149-165, adjust_contrast_image :188-205,
# _rgb_to_hsv :300-337, _hsv_to_rgb :340-367, adjust_hue_image :370-395.
# Python-float factors reach the fp32 kernels as float32(factor); `alpha = 1.0 - ratio` is formed in double first.
JITTER_BRIGHTNESS, JITTER_CONTRAST, JITTER_SATURATION, JITTER_HUE = 0, 1, 2, 3


def _gray(x):
    """_rgb_to_grayscale_image: r.mul(0.2989).add_(g, alpha=0.587).add_(b, alpha=0.114); ATen's add-with-alpha is
    a fused multiply-add in its vectorised loop."""
    f32 = np.float32
    r, g, b = x[:, 0], x[:, 1], x[:, 2]
    return _fma(b, f32(0.114), _fma(g, f32(0.587), _mul(r, f32(0.2989))))


def _blend(x, other, ratio: float):
    """_blend: image1.mul(ratio).add_(image2, alpha=1.0 - ratio).clamp_(0, 1)."""
    f32 = np.float32
    return np.clip(_fma(other, f32(1.0 - float(ratio)), _mul(x, f32(ratio))), f32(0), f32(1)).astype(np.float32)


def _adjust_hue(x, hue: float):
    f32 = np.float32
    r, g, b = x[:, 0], x[:, 1], x[:, 2]
    maxc = x.max(axis=1)
    minc = x.min(axis=1)
    eqc = maxc == minc
    cr = (maxc - minc).astype(np.float32)
    s = (cr / np.where(eqc, f32(1), maxc)).astype(np.float32)
    div = np.where(eqc, f32(1), cr).astype(np.float32)
    rc = ((maxc - r) / div).astype(np.float32)
    gc = ((maxc - g) / div).astype(np.float32)
    bc = ((maxc - b) / div).astype(np.float32)
    neq_r = maxc != r
    eq_g = maxc == g
    hg = ((rc + f32(2.0)).astype(np.float32) - bc).astype(np.float32) * (eq_g & neq_r)
    hr = (bc - gc).astype(np.float32) * (~neq_r)
    hb = ((gc + f32(4.0)).astype(np.float32) - rc).astype(np.float32) * (neq_r & ~eq_g)
    h = ((hr + hg).astype(np.float32) + hb).astype(np.float32)
    h = np.fmod((h * f32(1.0 / 6.0)).astype(np.float32) + f32(1.0), f32(1.0)).astype(np.float32)
    # h.add_(hue_factor).remainder_(1.0)
    h = np.remainder((h + f32(hue)).astype(np.float32), f32(1.0)).astype(np.float32)
    v = maxc
    h6 = (h * f32(6)).astype(np.float32)
    i = np.floor(h6)
    f = (h6 - i).astype(np.float32)
    i = np.remainder(i.astype(np.int32), 6)
    sxf = (s * f).astype(np.float32)
    oms = (f32(1.0) - s).astype(np.float32)
    q = np.clip(((f32(1.0) - sxf).astype(np.float32) * v).astype(np.float32), 0, 1)
    t = np.clip(((sxf + oms).astype(np.float32) * v).astype(np.float32), 0, 1)
    pp = np.clip((oms * v).astype(np.float32), 0, 1)
    vpqt = np.stack([v, pp, q, t], axis=0)  # (4, T, H, W)
    select = np.array([[0, 2, 1, 1, 3, 0], [3, 0, 0, 2, 1, 1], [1, 1, 3, 0, 0, 2]])
    out = np.stack([np.take_along_axis(vpqt, select[c][i][None], axis=0)[0] for c in range(3)], axis=1)
    return out.astype(np.float32)


def color_jitter(x01: np.ndarray, fn_idx, brightness: float, contrast: float, saturation: float, hue: float):
    """x01: (T,3,H,W) fp32 in [0,1] -> jittered clip, torchvision v2 ColorJitter.transform with explicit params."""
    f32 = np.float32
    x = x01.astype(np.float32)
    for fn in fn_idx:
        fn = int(fn)
        if fn == JITTER_BRIGHTNESS:  # image.mul(factor).clamp_(0, 1)
            x = np.clip(_mul(x, f32(brightness)), f32(0), f32(1)).astype(np.float32)
        elif fn == JITTER_CONTRAST:  # blend with the per-frame mean grey level
            mean = _gray(x).reshape(x.shape[0], -1).mean(axis=1, dtype=np.float64).astype(np.float32)
            x = _blend(x, mean[:, None, None, None], contrast)
        elif fn == JITTER_SATURATION:
            x = _blend(x, _gray(x)[:, None], saturation)
        elif fn == JITTER_HUE:
            x = _adjust_hue(x, hue)
        else:
            raise ValueError(fn)
    return x


def crop_resize_jitter_normalize(frames_u8: np.ndarray, box, fn_idx, brightness, contrast, saturation, hue,
                                 flip: bool = False, aten_path: str = "worker") -> np.ndarray:
    """dataset.py:141-152 -> [:166 hflip] -> :188-198 colour jitter -> :429 Normalize."""
    u8 = crop_resize_u8(frames_u8, box, 224, aten_path)
    x = u8.astype(np.float32) / np.float32(255.0)
    if flip:
        x = np.ascontiguousarray(x[..., ::-1])
    x = color_jitter(x, fn_idx, brightness, contrast, saturation, hue)
    return ((x - IMAGENET_MEAN[None, :, None, None]) / IMAGENET_STD[None, :, None, None]).astype(np.float32)


def to_nhwc4p_bf16_bits(x_nchw_f32: np.ndarray) -> np.ndarray:
    """fp32 (N,3,224,224) -> uint16 bf16 bit patterns in the library's NHWC4p layout (N,224,232,4), RNE rounding."""
    n = x_nchw_f32.shape[0]
    bits = x_nchw_f32.astype(np.float32).view(np.uint32)
    rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16).astype(np.uint16)
    out = np.zeros((n, 224, 232, 4), dtype=np.uint16)
    out[:, :, 4:228, 0:3] = rounded.transpose(0, 2, 3, 1)
    return out
