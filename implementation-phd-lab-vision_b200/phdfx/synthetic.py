"""Synthetic H36M-shaped clips for benchmarks, tests and `--synthetic` runs of the entry point.

Stands in for `Human36MPreprocessedClips` (src/dataset.py:210-437) when there is no dataset (the licensed H36M videos
and the video decoder are outside this drop-in): same index fields (`subject, action, cam, start, end`,
dataset.py:49-58) and the same per-clip outputs, except that frames stay uint8 and un-cropped, with the person box
next to them — crop / resize / normalise move to the GPU (K1)."""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from pathlib import Path
from typing import List, Tuple

import numpy as np
import torch
from torch.utils.data import Dataset

WEIGHT_SEED = 0
BN_SEED = 1


def seeded_backbone(weight_seed: int = WEIGHT_SEED, bn_seed: int = BN_SEED):
    """The reference's trunk construction (src/preprocess_resnet_features.py:207-209) with seeded random init (there
    is no network for IMAGENET1K_V2) and seeded non-trivial BatchNorm statistics, so BN folding is exercised.
    Benchmarks and `--synthetic` runs build their weights here; the oracle has its own copy of the same recipe
    (oracle/resnet50_ref.py) and tests/test_weights.py checks that the two agree tensor for tensor."""
    import torchvision

    from .weights import randomize_bn_

    torch.manual_seed(weight_seed)
    resnet = torchvision.models.resnet50(weights=None)
    backbone = torch.nn.Sequential(*list(resnet.children())[:-1]).eval()
    return randomize_bn_(backbone, bn_seed)


def seeded_frames(n: int, h: int, w: int, seed: int) -> np.ndarray:
    """uint8 (n,h,w,3) ~ U{0..255}; numpy PCG64, so the bytes are identical on every machine."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


def csrc_sha() -> str:
    """Content hash of the kernel sources + the C ABI header (the repo's .git does not travel to the GPU box): stamps
    profiler captures (tools/ncu_summarize.py) so bench.py can tell whether a committed capture describes the kernels
    that are running."""
    root = Path(__file__).resolve().parent.parent
    files = sorted((root / "csrc").glob("*.cu")) + sorted((root / "csrc").glob("*.cuh"))
    files.append(root.parent / "include" / "phdfx.h")
    h = hashlib.sha256()
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


@dataclass
class ClipIndex:
    subject: int
    action: str
    cam: str
    start: int
    end: int


class SyntheticH36MClips(Dataset):
    """len = n_clips; item = (frames uint8 (T,H,W,3), joints3d (T,17,3) mm, joints2d (T,17,2) px in the 224 crop,
    K (3,3), box int64 (top,left,h,w)).  Everything is a pure function of (seed, clip index).

    fast=True (throughput runs): the 6 MB of random bytes per clip come from ONE seeded base clip, rotated by a
    clip-dependent offset and XOR-ed with a clip-dependent byte (a memcpy instead of 18 ms of PRNG per clip), so the
    generator does not bound a multi-GPU run; every clip still has its own pixels, deterministically."""

    def __init__(self, n_clips: int, seq_len: int = 40, height: int = 224, width: int = 224,
                 subjects: Tuple[int, ...] = (1, 6, 7, 8), seed: int = 0, box_side: int = 0, fast: bool = False):
        self.n_clips, self.seq_len, self.h, self.w, self.seed = n_clips, seq_len, height, width, seed
        self.fast = fast
        self._base = None
        side = box_side if box_side > 0 else min(height, width)
        self.side = min(side, height, width)
        self.index: List[ClipIndex] = []
        for i in range(n_clips):
            self.index.append(ClipIndex(subject=int(subjects[i % len(subjects)]), action=f"Action_{(i // 7) % 15}",
                                        cam=f"cam_{i % 4}", start=5 * i, end=5 * i + seq_len))

    def __len__(self):
        return self.n_clips

    def box(self, i: int) -> torch.Tensor:
        rng = np.random.default_rng((self.seed, i, 1))
        top = int(rng.integers(0, self.h - self.side + 1))
        left = int(rng.integers(0, self.w - self.side + 1))
        return torch.tensor([top, left, self.side, self.side], dtype=torch.int64)

    def annotations(self, i: int):
        rng = np.random.default_rng((self.seed, i, 2))
        j3 = torch.from_numpy(rng.normal(0.0, 400.0, size=(self.seq_len, 17, 3)).astype(np.float32))
        j3[..., 2] += 4000.0
        j2 = torch.from_numpy(rng.uniform(0.0, 224.0, size=(self.seq_len, 17, 2)).astype(np.float32))
        K = torch.tensor([[500.0, 0.0, 112.0], [0.0, 500.0, 112.0], [0.0, 0.0, 1.0]])
        return j3, j2, K

    def frames(self, i: int) -> torch.Tensor:
        shape = (self.seq_len, self.h, self.w, 3)
        if self.fast:
            if self._base is None:  # per process (DataLoader workers each build their own copy)
                rng = np.random.default_rng((self.seed, 0, 3))
                self._base = rng.integers(0, 256, size=int(np.prod(shape)), dtype=np.uint8)
            out = np.roll(self._base, 7919 * (i + 1))
            out ^= np.uint8((37 * i + 11) & 0xFF)
            return torch.from_numpy(out.reshape(shape))
        rng = np.random.default_rng((self.seed, i, 0))
        return torch.from_numpy(rng.integers(0, 256, size=shape, dtype=np.uint8))

    def fill(self, i: int, out: torch.Tensor) -> None:
        """Write clip i's frames straight into `out` (uint8 [T,H,W,3], e.g. a slice of a pinned batch buffer): same
        bytes as frames(i).  fast=True does it in one pass (rotation and XOR fused into the copy; numpy releases the GIL,
        so several host threads fill different clips at memory speed)."""
        dst = out.numpy().reshape(-1)
        if not self.fast:
            dst[:] = self.frames(i).numpy().reshape(-1)
            return
        if self._base is None:
            self.frames(0)  # builds the base clip
        base, n = self._base, self._base.size
        k = (7919 * (i + 1)) % n
        c = np.uint8((37 * i + 11) & 0xFF)
        np.bitwise_xor(base[n - k:], c, out=dst[:k])  # np.roll(base, k)[:k] == base[n-k:]
        np.bitwise_xor(base[:n - k], c, out=dst[k:])

    def __getitem__(self, i: int):
        j3, j2, K = self.annotations(i)
        return self.frames(i), j3, j2, K, self.box(i)
