"""phdfx — B200-native ResNet-50 frame-feature extractor (host side of libphdfx.so).

Drop-in for the hot path of ferreiraluisa/implementation-phd-lab-vision:
`src/preprocess_resnet_features.py` (backbone forward, :207-209,:296) + `src/dataset.py:141-152,242-245`
(crop / resize / normalise).  See DESIGN.md and INTEGRATION.md.
"""
from ._lib import EXPORTS, FEAT_DIM, IMG, IN_CPAD, IN_LPAD, IN_WPAD, LIB_PATH, LayerDesc, load  # noqa: F401
from .weights import Plan, build_plan, fold_conv_bn, pack_conv, pack_stem, pack_stem_pool, randomize_bn_  # noqa: F401
from .backbone import B200Backbone, ExtractGraph, jitter_params  # noqa: F401
from .stream import StreamingExtractor  # noqa: F401
from .synthetic import csrc_sha, seeded_backbone, seeded_frames  # noqa: F401

__all__ = ["B200Backbone", "ExtractGraph", "jitter_params", "StreamingExtractor", "csrc_sha", "seeded_backbone", "seeded_frames", "build_plan", "fold_conv_bn", "pack_conv", "pack_stem", "pack_stem_pool", "randomize_bn_", "Plan", "load",
           "LayerDesc", "EXPORTS", "LIB_PATH", "FEAT_DIM", "IMG", "IN_WPAD", "IN_LPAD", "IN_CPAD"]
