"""C-ABI surface: the library builds, loads, exports every symbol include/phdfx.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest
import torch

import phdfx
from phdfx import _lib


@pytest.fixture(scope="module")
def lib(repo_root):
    if not os.path.exists(phdfx.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return phdfx.load()


def test_header_symbols_exported(lib, repo_root):
    hdr = open(os.path.join(repo_root, "include", "phdfx.h")).read()
    declared = sorted(set(re.findall(r"\b(phdfx_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/phdfx.h but not exported by libphdfx.so"
    assert sorted(phdfx.EXPORTS) == declared


def test_version(lib, repo_root):
    hdr = open(os.path.join(repo_root, "include", "phdfx.h")).read()
    ver = int(re.search(r"#define PHDFX_VERSION (\d+)", hdr).group(1))
    assert lib.phdfx_version() == ver


def test_layer_desc_layout_matches_header():
    # 18 x int32 + 2 x int64, no padding surprises
    assert C.sizeof(_lib.LayerDesc) == 18 * 4 + 2 * 8
    assert _lib.LayerDesc.w_off.offset == 72 and _lib.LayerDesc.b_off.offset == 80


def test_geometry_constants_match_header(repo_root):
    hdr = open(os.path.join(repo_root, "include", "phdfx.h")).read()
    consts = dict(re.findall(r"#define PHDFX_(IMG|IN_WPAD|IN_LPAD|IN_CPAD|FEAT_DIM) (\d+)", hdr))
    assert int(consts["IMG"]) == phdfx.IMG and int(consts["IN_WPAD"]) == phdfx.IN_WPAD
    assert int(consts["IN_LPAD"]) == phdfx.IN_LPAD and int(consts["IN_CPAD"]) == phdfx.IN_CPAD
    assert int(consts["FEAT_DIM"]) == phdfx.FEAT_DIM


def test_bad_arguments_are_rejected(lib):
    assert lib.phdfx_create(None, 0, 16) < 0
    h = C.c_void_p()
    assert lib.phdfx_create(C.byref(h), 0, 0) < 0  # max_frames must be >= 1
    assert b"max_frames" in lib.phdfx_last_error(None)
    assert lib.phdfx_layer_count(None) == 0
    assert lib.phdfx_destroy(None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    """Unlike the reference (src/preprocess_resnet_features.py:157-161) this backend refuses to run without CUDA."""
    h = C.c_void_p()
    rc = lib.phdfx_create(C.byref(h), 0, 16)
    assert rc < 0 and not h.value
    assert b"no CPU fallback" in lib.phdfx_last_error(None)
    import torchvision

    bb = torch.nn.Sequential(*list(torchvision.models.resnet50(weights=None).children())[:-1]).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        phdfx.B200Backbone(bb)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("PHDFX_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="not found"):
        _lib.load()
