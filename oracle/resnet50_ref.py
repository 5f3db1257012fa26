"""ORACLE (test infrastructure): ctypes wrapper around oracle/resnet50_ref.c + deterministic test-data builders.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
LIB = _DIR / "build" / "libresnet50_ref.so"


def build(force: bool = False) -> Path:
    """Compile the C oracle with oracle/Makefile (gcc only)."""
    if force or not LIB.exists() or LIB.stat().st_mtime < (_DIR / "resnet50_ref.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(str(LIB))
        lib.ref_resnet50_features.restype = C.c_int
        lib.ref_resnet50_features.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                              C.c_void_p, C.POINTER(C.c_void_p)]
        lib.ref_num_params.restype = C.c_int
        lib.ref_num_threads.restype = C.c_int
        lib.ref_conv2d.restype = None
        lib.ref_conv2d.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]
        _lib = lib
    return _lib


def num_threads() -> int:
    return int(_load().ref_num_threads())


def param_list(backbone) -> list:
    """fp32 numpy arrays in the order ref_resnet50_features expects (state_dict order of the trunk).

    `backbone` is the reference's nn.Sequential (src/preprocess_resnet_features.py:207-208) or a torchvision ResNet."""
    import torch.nn as nn

    if isinstance(backbone, nn.Sequential):
        mods = list(backbone.children())
        conv1, bn1, stages = mods[0], mods[1], mods[4:8]
    else:
        conv1, bn1 = backbone.conv1, backbone.bn1
        stages = [backbone.layer1, backbone.layer2, backbone.layer3, backbone.layer4]

    def arr(t):
        return np.ascontiguousarray(t.detach().cpu().numpy().astype(np.float32))

    def conv_bn(conv, bn):
        return [arr(conv.weight), arr(bn.weight), arr(bn.bias), arr(bn.running_mean), arr(bn.running_var)]

    out = conv_bn(conv1, bn1)
    for stage in stages:
        for blk in stage:
            out += conv_bn(blk.conv1, blk.bn1) + conv_bn(blk.conv2, blk.bn2) + conv_bn(blk.conv3, blk.bn3)
            if blk.downsample is not None:
                out += conv_bn(blk.downsample[0], blk.downsample[1])
    return out


def features(x_nchw: np.ndarray, params: list, taps: bool = False):
    """x: fp32 (N,3,H,W) ImageNet-normalised -> fp32 (N,2048).  taps=True also returns the activations after the
    stem+maxpool and after each stage (NCHW)."""
    lib = _load()
    x = np.ascontiguousarray(x_nchw, dtype=np.float32)
    n, c, h, w = x.shape
    assert c == 3
    assert len(params) == lib.ref_num_params(), (len(params), lib.ref_num_params())
    ptrs = (C.c_void_p * len(params))(*[p.ctypes.data for p in params])
    feats = np.empty((n, 2048), dtype=np.float32)
    tap_arrays = None
    tap_ptrs = None
    if taps:
        assert h == 224 and w == 224
        shapes = [(n, 64, 56, 56), (n, 256, 56, 56), (n, 512, 28, 28), (n, 1024, 14, 14), (n, 2048, 7, 7)]
        tap_arrays = [np.empty(s, dtype=np.float32) for s in shapes]
        tap_ptrs = (C.c_void_p * 5)(*[t.ctypes.data for t in tap_arrays])
    rc = lib.ref_resnet50_features(x.ctypes.data, n, h, w, ptrs, feats.ctypes.data, tap_ptrs)
    if rc != 0:
        raise MemoryError("oracle: allocation failed")
    return (feats, tap_arrays) if taps else feats


# ---------------------------------------------------------------------------------------------------------------
# deterministic test data (shared by oracle/make_golden.py, tests/ and bench.py so fixtures can be regenerated)
# ---------------------------------------------------------------------------------------------------------------
WEIGHT_SEED = 0
BN_SEED = 1


def seeded_backbone(weight_seed: int = WEIGHT_SEED, bn_seed: int = BN_SEED):
    """The reference's trunk construction (:207-209) with seeded random init (no network for IMAGENET1K_V2) and
    seeded non-trivial BatchNorm statistics so that BN folding is actually exercised."""
    import torch
    import torchvision

    torch.manual_seed(weight_seed)
    resnet = torchvision.models.resnet50(weights=None)
    backbone = torch.nn.Sequential(*list(resnet.children())[:-1]).eval()
    g = torch.Generator().manual_seed(bn_seed)
    for m in backbone.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            nf = m.num_features
            with torch.no_grad():
                m.running_mean.copy_(0.1 * torch.randn(nf, generator=g))
                m.running_var.copy_(0.5 + torch.rand(nf, generator=g))
                m.weight.copy_(0.6 + 0.5 * torch.rand(nf, generator=g))
                m.bias.copy_(0.1 * torch.randn(nf, generator=g))
    return backbone


def seeded_frames(n: int, h: int, w: int, seed: int) -> np.ndarray:
    """uint8 (n,h,w,3) ~ U{0..255}; numpy PCG64 so the bytes are identical on every machine."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
