"""On-disk feature layout — the contract between the extractor and everything downstream.

Produces exactly what the reference extractor writes (src/preprocess_resnet_features.py:80-131, 344-417) and what
its reader consumes (src/dataset_features.py:16-27, 50-59, 107-121; src/samplers.py:22-25):

  out/shard_{sid:05d}.pt   legacy (non-zip) torch pickle of
       {"feats": (rows,T,2048) fp32|fp16, "joints3d": (rows,T,17,3), "joints2d": (rows,T,17,2), "K": (rows,3,3),
        "meta": [dict]*rows, "n_vars": int}          rows = clips_in_shard * n_vars, variants of a clip contiguous
  out/index.pt             {"clips": [{shard_id,row,subject,action,cam,start,end}], "n_shards", "n_clips",
        "n_variants", "aug_names", "seq_len", "frame_skip", "feat_dtype", "variants_grouped", "shuffle_seed",
        "shuffle_pool"}

Shuffle semantics are the reference's: clips are pooled in arrival order; whenever the pool reaches `shuffle_pool`
clips, carry-over + pool is shuffled by ONE `random.Random(shuffle_seed)` stream and cut into full shards, the
remainder carried over; the final flush shuffles once more and writes full shards plus one partial shard.

What differs is the mechanics: rows are written once into a preallocated shard tensor instead of
`torch.stack` over up to 2048 small tensors, and files are written by a bounded background thread.

Multi-GPU runs do not funnel features through one writer at all: which clip lands in which row of
which shard is a pure function of (n_clips, shard_size, shuffle_pool, shuffle_seed) — `plan_shards` replays the
reference's pooling / shuffling on clip NUMBERS — so every rank extracts exactly the clips of the shards it owns and
writes those files itself (`assemble_shard`), and rank 0 writes `index.pt` from the plan (`index_from_plan`).  The
output holds exactly what the single-process writer produces (same index, same tensors, same row order).
"""
from __future__ import annotations

import queue
import random
import threading
from pathlib import Path
from typing import List, Optional, Sequence

import torch

AUG_NAMES = ["orig", "cjitter", "hflip", "trev"]  # src/preprocess_resnet_features.py:27


class AsyncShardWriter:
    """Bounded background writer: torch.save(..., _use_new_zipfile_serialization=False) off the hot loop
    (same file format as the reference's AsyncFileWriter, :29-57; errors are re-raised instead of lost)."""

    def __init__(self, max_queue: int = 8):
        self._q: "queue.Queue" = queue.Queue(maxsize=max_queue)
        self._err: Optional[BaseException] = None
        self.count = 0
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            try:
                if item is None:
                    return
                obj, path = item
                torch.save(obj, path, _use_new_zipfile_serialization=False)
            except BaseException as e:  # noqa: BLE001
                self._err = e
            finally:
                self._q.task_done()

    def save(self, obj, path):
        if self._err:
            raise RuntimeError("shard writer failed") from self._err
        self._q.put((obj, str(path)))
        self.count += 1

    def wait(self):
        self._q.join()
        if self._err:
            raise RuntimeError("shard writer failed") from self._err

    def stop(self):
        self._q.put(None)
        self._t.join()
        if self._err:
            raise RuntimeError("shard writer failed") from self._err


class ClipRecord:
    """One clip with its n_vars variants (a 'group' in the reference, :299-323)."""

    __slots__ = ("feats", "joints3d", "joints2d", "K", "metas")

    def __init__(self, feats: Sequence[torch.Tensor], joints3d: Sequence[torch.Tensor],
                 joints2d: Sequence[torch.Tensor], K: Sequence[torch.Tensor], metas: Sequence[dict]):
        self.feats, self.joints3d, self.joints2d, self.K, self.metas = feats, joints3d, joints2d, K, metas


def assemble_shard(groups: Sequence[ClipRecord], n_vars: int) -> dict:
    """The shard dict of `groups` (one ClipRecord per clip, variants contiguous): rows filled in place."""
    rows = len(groups) * n_vars
    g0 = groups[0]
    feats = torch.empty((rows,) + tuple(g0.feats[0].shape), dtype=g0.feats[0].dtype)
    j3 = torch.empty((rows,) + tuple(g0.joints3d[0].shape), dtype=g0.joints3d[0].dtype)
    j2 = torch.empty((rows,) + tuple(g0.joints2d[0].shape), dtype=g0.joints2d[0].dtype)
    Ks = torch.empty((rows,) + tuple(g0.K[0].shape), dtype=g0.K[0].dtype)
    metas = []
    for i, g in enumerate(groups):
        base = i * n_vars
        for v in range(n_vars):
            feats[base + v] = g.feats[v]
            j3[base + v] = g.joints3d[v]
            j2[base + v] = g.joints2d[v]
            Ks[base + v] = g.K[v]
            metas.append(g.metas[v])
    return {"feats": feats, "joints3d": j3, "joints2d": j2, "K": Ks, "meta": metas, "n_vars": n_vars}


def shard_path(out_root, shard_id: int) -> Path:
    return Path(out_root) / f"shard_{shard_id:05d}.pt"


def plan_shards(n_clips: int, shard_size: int, shuffle_pool: int, shuffle_seed: int) -> List[List[int]]:
    """Clip numbers (dataset order) of every shard, in row order: the reference's pooling and shuffling
    (src/preprocess_resnet_features.py:94-131, 269, 325-396) replayed on integers.  `random.Random.shuffle` draws
    depend only on the list length, so this is exactly the permutation the streaming writer applies to clip records."""
    rng = random.Random(shuffle_seed)
    pool: List[int] = []
    carry: List[int] = []
    shards: List[List[int]] = []

    def flush(final: bool):
        nonlocal pool, carry
        combined = carry + pool
        rng.shuffle(combined)
        n_full = len(combined) // shard_size
        for s in range(n_full):
            shards.append(combined[s * shard_size:(s + 1) * shard_size])
        rest = combined[n_full * shard_size:]
        pool = []
        if final:
            if rest:
                shards.append(rest)
            carry = []
        else:
            carry = rest

    for i in range(n_clips):
        pool.append(i)
        if len(pool) >= shuffle_pool:
            flush(False)
    flush(True)
    return shards


def index_from_plan(plan: Sequence[Sequence[int]], clip_meta, n_vars: int, seq_len: int, frame_skip: int,
                    save_fp16: bool, augment: bool, shuffle_seed: int, shuffle_pool: int) -> dict:
    """`index.pt` content for a planned run.  clip_meta(i) -> object/dict with subject, action, cam, start, end of
    clip i (the dataset's own index entry)."""
    clips = []
    for sid, ids in enumerate(plan):
        for row, i in enumerate(ids):
            m = clip_meta(i)
            get = (lambda k: m[k]) if isinstance(m, dict) else (lambda k: getattr(m, k))
            clips.append({"shard_id": sid, "row": row * n_vars, "subject": get("subject"), "action": get("action"),
                          "cam": get("cam"), "start": get("start"), "end": get("end")})
    return {
        "clips": clips, "n_shards": len(plan), "n_clips": sum(len(ids) for ids in plan), "n_variants": n_vars,
        "aug_names": list(AUG_NAMES) if augment else ["orig"], "seq_len": seq_len, "frame_skip": frame_skip,
        "feat_dtype": "float16" if save_fp16 else "float32", "variants_grouped": True, "shuffle_seed": shuffle_seed,
        "shuffle_pool": shuffle_pool,
    }


class ShardWriter:
    def __init__(self, out_root, n_vars: int, shard_size: int = 512, shuffle_pool: int = 8192,
                 shuffle_seed: int = 123, writer: Optional[AsyncShardWriter] = None):
        self.out_root = Path(out_root)
        self.out_root.mkdir(parents=True, exist_ok=True)
        self.n_vars = int(n_vars)
        self.shard_size = int(shard_size)
        self.shuffle_pool = int(shuffle_pool)
        self.shuffle_seed = int(shuffle_seed)
        self.rng = random.Random(shuffle_seed)  # :269
        self.writer = writer or AsyncShardWriter()
        self._own_writer = writer is None
        self.pool: List[ClipRecord] = []
        self.carry: List[ClipRecord] = []
        self.clip_index: List[dict] = []
        self.shard_id = 0
        self.n_clips = 0

    # ---- write one shard: rows filled in place -----------------------------------------------------------------
    def _write_shard(self, groups: List[ClipRecord]):
        for i, g in enumerate(groups):
            m0 = g.metas[0]
            self.clip_index.append({"shard_id": self.shard_id, "row": i * self.n_vars, "subject": m0["subject"],
                                    "action": m0["action"], "cam": m0["cam"], "start": m0["start"],
                                    "end": m0["end"]})
        self.writer.save(assemble_shard(groups, self.n_vars), shard_path(self.out_root, self.shard_id))
        self.shard_id += 1

    def _flush(self, final: bool):
        combined = self.carry + self.pool
        self.rng.shuffle(combined)
        n_full = len(combined) // self.shard_size
        for s in range(n_full):
            self._write_shard(combined[s * self.shard_size:(s + 1) * self.shard_size])
        rest = combined[n_full * self.shard_size:]
        self.pool = []
        if final:
            if rest:
                self._write_shard(rest)
            self.carry = []
        else:
            self.carry = rest

    def add(self, rec: ClipRecord):
        if len(rec.feats) != self.n_vars:
            raise ValueError(f"clip has {len(rec.feats)} variants, writer expects {self.n_vars}")
        self.pool.append(rec)
        self.n_clips += 1
        if len(self.pool) >= self.shuffle_pool:
            self._flush(final=False)

    def finish(self, seq_len: int, frame_skip: int, save_fp16: bool, augment: bool) -> dict:
        self._flush(final=True)
        self.writer.wait()
        if self._own_writer:
            self.writer.stop()
        index = {
            "clips": self.clip_index, "n_shards": self.shard_id, "n_clips": self.n_clips,
            "n_variants": self.n_vars, "aug_names": list(AUG_NAMES) if augment else ["orig"], "seq_len": seq_len,
            "frame_skip": frame_skip, "feat_dtype": "float16" if save_fp16 else "float32",
            "variants_grouped": True, "shuffle_seed": self.shuffle_seed, "shuffle_pool": self.shuffle_pool,
        }
        torch.save(index, self.out_root / "index.pt")  # zip format, like the reference (:403-417)
        return index
