"""Pins oracle/preprocess_ref.py to the golden vectors produced by the REFERENCE's own functions
(src/dataset.py:141-152,242-245, via oracle/make_golden.py): bit-exact uint8 and bit-exact fp32."""
import os

import numpy as np
import pytest

import preprocess_ref as P
import resnet50_ref as R


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess_golden.npz"))


def _cases(golden):
    i = 0
    while f"case{i}_meta" in golden:
        n, H, W, top, left, hh, ww, seed = (int(v) for v in golden[f"case{i}_meta"])
        yield i, R.seeded_frames(n, H, W, seed), (top, left, hh, ww)
        i += 1


def test_resize_bit_exact_worker_path(golden):
    n_cases = 0
    for i, frames, box in _cases(golden):
        got = P.crop_resize_u8(frames, box, aten_path="worker")
        assert np.array_equal(got, golden[f"case{i}_u8_worker"]), f"case {i} box {box}"
        n_cases += 1
    assert n_cases >= 6


def test_resize_bit_exact_generic_path(golden):
    for i, frames, box in _cases(golden):
        got = P.crop_resize_u8(frames, box, aten_path="generic")
        assert np.array_equal(got, golden[f"case{i}_u8_generic"]), f"case {i} box {box}"


def test_the_two_aten_paths_differ_by_at_most_one_lsb(golden):
    for i, _, _ in _cases(golden):
        a = golden[f"case{i}_u8_worker"].astype(np.int16)
        b = golden[f"case{i}_u8_generic"].astype(np.int16)
        assert np.abs(a - b).max() <= 1
        assert (a != b).mean() < 2e-3


def test_normalise_bit_exact(golden):
    _, frames, box = next(_cases(golden))
    got = P.crop_resize_normalize(frames, box)
    ref = golden["case0_norm_worker"]
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_identity_crop_is_untouched(golden):
    frames = R.seeded_frames(2, 224, 224, 99)
    got = P.crop_resize_u8(frames, (0, 0, 224, 224))
    assert np.array_equal(got, frames.transpose(0, 3, 1, 2))


def test_nhwc4p_layout():
    x = np.random.default_rng(0).standard_normal((2, 3, 224, 224)).astype(np.float32)
    out = P.to_nhwc4p_bf16_bits(x)
    assert out.shape == (2, 224, 232, 4) and out.dtype == np.uint16
    assert not out[:, :, :4].any() and not out[:, :, 228:].any() and not out[..., 3].any()
    import torch

    ref = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(out[:, :, 4:228, :3], ref.transpose(0, 2, 3, 1))


# ------------------------------------------------------------------------------------------------ colour jitter (8f N1)
JITTER_TOL = 1e-6  # fp32 rounding of the contrast op's mean (torch sums in fp32, the oracle in fp64) and FMA choices


def _tv_jitter(x, order, b, c, s, h):
    """torchvision's own kernels in the given order (what v2 ColorJitter.transform does, _color.py:156-173)."""
    import torch
    from torchvision.transforms.v2 import functional as F

    t = torch.from_numpy(x.copy())
    for fn in order:
        t = (F.adjust_brightness(t, b), F.adjust_contrast(t, c), F.adjust_saturation(t, s), F.adjust_hue(t, h))[fn]
    return t.numpy()


@pytest.mark.parametrize("order", [(0, 1, 2, 3), (3, 2, 1, 0), (1, 0, 3, 2), (2, 3, 0, 1), (1, 2, 3, 0), (3, 0, 1, 2)])
def test_color_jitter_vs_torchvision(order):
    """oracle.color_jitter == torchvision v2 adjust_* applied in the same order, on uint8-derived clips incl. grey
    pixels (max == min: the hue op's guarded divisions) and saturated ones."""
    g = np.random.default_rng(sum(order[i] * 4 ** i for i in range(4)))
    x = (g.integers(0, 256, (3, 3, 40, 44)).astype(np.float32) / np.float32(255.0))
    x[:, :, :6] = x[:, :1, :6]      # grey rows
    x[0, :, 6:9] = 0.0              # black
    x[1, :, 9:12] = 1.0             # white
    b, c, s, h = (float(v) for v in (g.uniform(0.7, 1.3), g.uniform(0.7, 1.3), g.uniform(0.8, 1.2),
                                     g.uniform(-0.05, 0.05)))
    got = P.color_jitter(x, order, b, c, s, h)
    ref = _tv_jitter(x, order, b, c, s, h)
    assert got.dtype == np.float32 and got.min() >= 0.0 and got.max() <= 1.0
    assert np.abs(got - ref).max() <= JITTER_TOL


def test_color_jitter_vs_the_references_function():
    """The reference's `_aug_color_jitter` (src/dataset.py:188-198) under a fixed torch seed, against the oracle fed
    with the parameters the same seed draws (ColorJitter.make_params consumes the RNG exactly like the call does)."""
    from conftest import REFERENCE_SRC

    if not os.path.isdir(REFERENCE_SRC):
        pytest.skip("reference sources only exist in the build container")
    import sys

    import torch
    import torchvision.io as tio

    if not hasattr(tio, "VideoReader"):
        tio.VideoReader = None  # src/dataset.py:14 imports a name newer torchvision dropped
    sys.path.insert(0, REFERENCE_SRC)
    import dataset as ref_dataset
    from torchvision.transforms import v2 as T2

    frames = R.seeded_frames(4, 300, 280, 17)
    x = P.crop_resize_u8(frames, (10, 20, 231, 231)).astype(np.float32) / np.float32(255.0)
    for seed in (0, 1, 2, 3):
        torch.manual_seed(seed)
        ref = ref_dataset._aug_color_jitter(torch.from_numpy(x.copy())).numpy()
        torch.manual_seed(seed)
        prm = T2.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.2, hue=0.05).make_params([])
        got = P.color_jitter(x, [int(v) for v in prm["fn_idx"]], prm["brightness_factor"], prm["contrast_factor"],
                             prm["saturation_factor"], prm["hue_factor"])
        assert np.abs(got - ref).max() <= JITTER_TOL, seed
