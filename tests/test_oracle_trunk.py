"""Pins oracle/resnet50_ref.c to golden features computed by the reference's own arithmetic
(torchvision ResNet-50, CPU fp32, constructed exactly as src/preprocess_resnet_features.py:207-209)."""
import os

import numpy as np
import pytest

import preprocess_ref as P
import resnet50_ref as R

TOL = 1e-4  # fp32 vs fp32, different summation order; normalised by the per-frame max |feature|


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "trunk_golden.npz"))


@pytest.fixture(scope="module")
def params():
    return R.param_list(R.seeded_backbone())


def test_param_count(params):
    assert len(params) == 5 * 53  # 53 conv+BN pairs
    assert sum(p.size for p in params[::5]) == 23_454_912  # conv weights


def test_features_match_torchvision_golden(golden, params):
    n, H, W, top, left, hh, ww, seed = (int(v) for v in golden["meta"][:8])
    frames = R.seeded_frames(n, H, W, seed)[:2]
    x = P.crop_resize_normalize(frames, (top, left, hh, ww))
    feats, taps = R.features(x, params, taps=True)
    ref = golden["feats"][:2]
    err = np.abs(feats - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() < TOL, err
    # intermediate activations: after stem+maxpool (module index 3) and after layer4 (index 7)
    for tap_i, mod_i in ((0, 3), (1, 4), (2, 5), (3, 6), (4, 7)):
        ref_plane = golden[f"tap{mod_i}_frame0_ch0"]
        got_plane = taps[tap_i][0, 0]
        assert np.abs(got_plane - ref_plane).max() <= TOL * max(1.0, float(golden[f"tap{mod_i}_max"]))


def test_identity_size_input(golden, params):
    frames = R.seeded_frames(4, 224, 224, 2)[:1]
    x = P.crop_resize_normalize(frames, (0, 0, 224, 224))
    feats = R.features(x, params)
    ref = golden["feats_identity"][:1]
    assert (np.abs(feats - ref).max() / np.abs(ref).max()) < TOL


def test_conv2d_edge_cases():
    """stride-2 / padded conv against a direct numpy evaluation on a tiny case."""
    import ctypes as C

    rng = np.random.default_rng(0)
    for (Cn, H, K, Rr, stride, pad) in [(3, 9, 4, 3, 2, 1), (2, 7, 3, 1, 2, 0), (3, 11, 2, 7, 2, 3), (4, 5, 2, 3, 1, 1)]:
        x = rng.standard_normal((2, Cn, H, H)).astype(np.float32)
        w = rng.standard_normal((K, Cn, Rr, Rr)).astype(np.float32)
        Pn = (H + 2 * pad - Rr) // stride + 1
        y = np.empty((2, K, Pn, Pn), dtype=np.float32)
        R._load().ref_conv2d(x.ctypes.data, 2, Cn, H, H, w.ctypes.data, K, Rr, Rr, stride, pad, y.ctypes.data)
        xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
        ref = np.zeros_like(y, dtype=np.float64)
        for p in range(Pn):
            for q in range(Pn):
                patch = xp[:, :, p * stride:p * stride + Rr, q * stride:q * stride + Rr]
                ref[:, :, p, q] = np.einsum("ncrs,kcrs->nk", patch.astype(np.float64), w.astype(np.float64))
        assert np.abs(y - ref).max() < 1e-4
