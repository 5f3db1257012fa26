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
