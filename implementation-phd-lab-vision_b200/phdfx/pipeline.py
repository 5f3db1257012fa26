"""The extractor's driver loop at B200 rates — what replaces the reference's per-batch loop and clip grouping
(src/preprocess_resnet_features.py:273-330) and its shard assembly (:80-131) when the engine delivers ~10^5 frames/s.

The reference handles one clip at a time in Python: per variant a blocking `.cpu()` (:297), per clip a dict of small
tensors (:299-323), per shard a `torch.stack` over up to 2048 of them (:80-91).  Here nothing on the hot loop is per
clip:

  * which clip lands in which row of which shard is known up front (phdfx.shards.plan_shards), so clips are fetched in
    shard-row order and a batch of B clips IS rows [r0, r0 + B*n_vars) of its shard;
  * the B*T frames of a batch go up in one pinned H2D copy; the engine runs the 3 computed variants (orig, colour
    jitter, h-flip) on them; the 4th (time reversal) is a flip of the first (frames are independent);
    the [B, n_vars, T, 2048] block is put together on the device and ONE D2H copy lands it in the shard's preallocated
    pinned feature tensor — the tensor torch.save later writes, no stacking, no second copy;
  * annotations are transformed per batch with vectorised ops (augment_annotations_batch) straight into the shard's
    preallocated tensors;
  * upload of batch i+1, compute of batch i and download of batch i-1 run on three streams; the host thread only
    enqueues.  A finished shard goes to a writer thread together with the CUDA event of its last download; shard
    buffers come from a small pool, so the time the loop waits for a free buffer is exactly the time the writer is
    the bottleneck (reported as `writer_wait_s`).

Works on CPU tensors too (no streams, plain copies): the `--backend torch` comparison arm and the CPU tests run the
same code.
"""
from __future__ import annotations

import queue
import threading
import time
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, List, Optional, Sequence

import torch

from .shards import AUG_NAMES, AsyncShardWriter, shard_path

H36M_FLIP_PAIRS = ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))  # left/right joints (src/dataset.py:39-46)


def augment_annotations_batch(j3: torch.Tensor, j2: torch.Tensor, K: torch.Tensor, width: int = 224):
    """Annotation side of the four variants for a whole batch (src/dataset.py:158-207, 411-426):
    j3 [B,T,J,3], j2 [B,T,J,2], K [B,3,3] -> ([B,4,T,J,3], [B,4,T,J,2], [B,4,3,3]) in AUG_NAMES order
    (orig, cjitter, hflip, trev).  hflip mirrors x, negates camera-space x and swaps the left/right joints; trev
    reverses time; colour jitter leaves annotations alone."""
    j3f, j2f, Kf = j3.clone(), j2.clone(), K.clone()
    j2f[..., 0] = width - j2f[..., 0]
    j3f[..., 0] = -j3f[..., 0]
    perm = list(range(j3.shape[2]))
    for l, r in H36M_FLIP_PAIRS:
        perm[l], perm[r] = r, l
    j2f, j3f = j2f[:, :, perm], j3f[:, :, perm]
    Kf[:, 0, 2] = width - Kf[:, 0, 2]
    return (torch.stack([j3, j3, j3f, torch.flip(j3, dims=[1])], dim=1),
            torch.stack([j2, j2, j2f, torch.flip(j2, dims=[1])], dim=1),
            torch.stack([K, K, Kf, K], dim=1))


@dataclass
class ClipBatch:
    """What a feeder hands over for a list of clip ids (in that order)."""
    ids: List[int]
    groups: list  # [(positions in the batch: LongTensor, frames uint8 [g,T,H,W,3], boxes int64 [g,4])], one per (H, W)
    j3: torch.Tensor  # [B,T,J,3]
    j2: torch.Tensor  # [B,T,J,2]
    K: torch.Tensor  # [B,3,3]
    boxes: torch.Tensor  # [B,4] int64
    copied: object = None  # set by the consumer: CUDA event after which the frames' host memory may be reused


def collate_clips(items) -> tuple:
    """DataLoader collate_fn (runs in the worker): stack per frame size, so the main process gets whole tensors that the
    loader's pin-memory thread can pin — not B separate clips to stack (and page-lock) on the hot loop."""
    by_shape = {}
    for pos, it in enumerate(items):
        by_shape.setdefault(tuple(it[0].shape[1:3]), []).append(pos)
    groups = [(torch.tensor(pos, dtype=torch.long), torch.stack([items[p][0] for p in pos]),
               torch.stack([items[p][4] for p in pos]).to(torch.int64)) for pos in by_shape.values()]
    return (groups, torch.stack([it[1] for it in items]), torch.stack([it[2] for it in items]),
            torch.stack([it[3] for it in items]), torch.stack([it[4] for it in items]).to(torch.int64))


class ClipFeeder:
    """Iterates ClipBatch objects for a fixed list of clip-id lists.  workers > 0: one persistent DataLoader decodes /
    generates batch i+1.. in worker processes while batch i is on the GPU (the reference's worker pool, :195-204,
    driven by an explicit batch list); collation and pinning happen off the main thread.  Datasets that can write a clip
    straight into a caller-provided tensor (`fill(i, out)`, e.g. the fast synthetic clips) are instead served by
    `threads` host threads filling pinned batch buffers directly — no worker -> shared memory -> pinned copy chain."""

    def __init__(self, ds, batches: Sequence[Sequence[int]], workers: int, pin: bool, threads: int = 0, depth: int = 3):
        self.ds, self.batches, self.pin = ds, [list(b) for b in batches if len(b)], pin
        self.workers, self.threads, self.depth = workers, threads, depth

    def __iter__(self):
        if self.threads > 0 and hasattr(self.ds, "fill"):
            return self._iter_threads()
        if self.workers > 0 and self.batches:
            from torch.utils.data import DataLoader

            loader = DataLoader(self.ds, batch_sampler=self.batches, num_workers=min(self.workers, len(self.batches)),
                                collate_fn=collate_clips, pin_memory=self.pin, prefetch_factor=2)
            return (ClipBatch(ids, *c) for ids, c in zip(self.batches, loader))
        return (ClipBatch(ids, *collate_clips([self.ds[i] for i in ids])) for ids in self.batches)

    def _iter_threads(self):
        from concurrent.futures import ThreadPoolExecutor

        if not self.batches:
            return
        T, H, W = self.ds.seq_len, self.ds.h, self.ds.w
        bmax = max(len(b) for b in self.batches)
        slots = []
        for _ in range(self.depth + 1):
            t = torch.empty(bmax, T, H, W, 3, dtype=torch.uint8)
            slots.append(t.pin_memory() if self.pin else t)
        pool = ThreadPoolExecutor(max_workers=self.threads)

        handed = {}  # slot -> the ClipBatch that last used it

        def build(k):
            ids = self.batches[k]
            sl = k % len(slots)
            prev = handed.pop(sl, None)
            if prev is not None and prev.copied is not None:
                prev.copied.synchronize()  # the slot's previous frames have been uploaded
            buf = slots[sl][:len(ids)]
            futs = [pool.submit(one, i, buf[j]) for j, i in enumerate(ids)]
            return k, ids, buf, futs

        def one(i, out):  # frames into the pinned slot, annotations alongside: all off the main thread
            self.ds.fill(i, out)
            return self.ds.annotations(i), self.ds.box(i)

        pending = [build(k) for k in range(min(self.depth, len(self.batches)))]
        nxt = len(pending)
        try:
            while pending:
                k, ids, buf, futs = pending.pop(0)
                res = [f.result() for f in futs]
                ann = [r[0] for r in res]
                boxes = torch.stack([r[1] for r in res]).to(torch.int64)
                cb = ClipBatch(ids, [(torch.arange(len(ids)), buf, boxes)], torch.stack([a[0] for a in ann]),
                               torch.stack([a[1] for a in ann]), torch.stack([a[2] for a in ann]), boxes)
                handed[k % len(slots)] = cb
                yield cb
                # the slot handed out `depth` batches ago is refilled only now: its H2D copy was enqueued long before
                if nxt < len(self.batches):
                    pending.append(build(nxt))
                    nxt += 1
        finally:
            pool.shutdown(wait=True)


@dataclass
class ShardBuffers:
    feats: torch.Tensor
    j3: torch.Tensor
    j2: torch.Tensor
    K: torch.Tensor
    metas: list = field(default_factory=list)


@dataclass
class PipelineStats:
    clips: int = 0
    shards: int = 0
    frames_computed: int = 0
    total_s: float = 0.0
    writer_wait_s: float = 0.0  # time the loop waited for a free shard buffer = the writer was the bottleneck
    feed_wait_s: float = 0.0    # time the loop waited for the next batch of frames = the feed was the bottleneck
    write_s: float = 0.0        # time the writer thread spent inside torch.save (overlapped with the loop)
    bytes_written: int = 0


class PlannedShardRun:
    """Extract the clips of the shards in `my_shards` (ids into `plan`) and write those shard files.

    extract(frames [N,H,W,3] uint8 on `device`, boxes [N,4] int32 on `device`, flip: bool, jitter: Optional[fp32 [N,12]
    on `device`]) -> fp32 [N,2048] on `device`;  jitter_rows(clip_ids) -> fp32 [B,12] (CPU) or None for no colour jitter
    support (then the variant must not be requested)."""

    def __init__(self, *, extract: Callable, device: torch.device, out_root, n_vars: int, seq_len: int,
                 feat_dtype: torch.dtype, augment: bool, jitter_rows: Optional[Callable], clip_meta: Callable,
                 n_buffers: int = 2, log: Callable = print):
        self.extract, self.device, self.out_root = extract, torch.device(device), Path(out_root)
        self.n_vars, self.T, self.feat_dtype, self.augment = n_vars, seq_len, feat_dtype, augment
        self.jitter_rows, self.clip_meta, self.log = jitter_rows, clip_meta, log
        self.cuda = self.device.type == "cuda"
        self.n_buffers = n_buffers
        self.stats = PipelineStats()
        self._free: "queue.Queue[ShardBuffers]" = queue.Queue()
        self._full_rows = 0
        if self.cuda:
            self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))

    # ---- shard buffers ------------------------------------------------------------------------------------------
    def _host(self, shape, dtype):
        t = torch.empty(shape, dtype=dtype)
        return t.pin_memory() if self.cuda else t

    def _alloc(self, rows: int, ann: ClipBatch) -> ShardBuffers:
        return ShardBuffers(self._host((rows, self.T, 2048), self.feat_dtype),
                            self._host((rows,) + tuple(ann.j3.shape[1:]), ann.j3.dtype),
                            self._host((rows,) + tuple(ann.j2.shape[1:]), ann.j2.dtype),
                            self._host((rows,) + tuple(ann.K.shape[1:]), ann.K.dtype))

    def _get_buffers(self, rows: int, ann: ClipBatch) -> ShardBuffers:
        """Full-size shards rotate through a pool of pinned buffers (torch.save writes a tensor's whole storage, so a
        buffer is only ever used for a shard of exactly its size); an odd-sized (last) shard gets its own."""
        if rows != self._full_rows and self._full_rows:
            return self._alloc(rows, ann)
        if not self._full_rows:
            self._full_rows = rows
            for _ in range(self.n_buffers):
                self._free.put(self._alloc(rows, ann))
        t0 = time.perf_counter()
        buf = self._free.get()
        self.stats.writer_wait_s += time.perf_counter() - t0
        if isinstance(buf, BaseException):
            raise RuntimeError("shard writer failed") from buf
        buf.metas = []
        return buf

    # ---- one batch ------------------------------------------------------------------------------------------------
    def _block(self, batch: ClipBatch, slot: int):
        """Device block [B, n_vars, T, 2048] (feat_dtype) for the batch."""
        B, T = len(batch.ids), self.T
        block = self._d_block(slot, B)
        jrows = None
        if self.augment:
            jrows = self.jitter_rows(batch.ids)  # [B,12] CPU
        for pos, frames, boxes in batch.groups:
            g, _, H, W, _ = frames.shape
            flat = frames.view(g * T, H, W, 3)
            bx = boxes.to(torch.int32).repeat_interleave(T, dim=0)
            jr = jrows[pos].repeat_interleave(T, dim=0) if jrows is not None else None
            if self.cuda:
                with torch.cuda.stream(self.s_in):
                    if self._ev_done[slot] is not None:
                        self.s_in.wait_event(self._ev_done[slot])  # the slot's previous batch has been consumed
                    d_fr = self._d_frames(slot, flat.numel()).view(g * T, H, W, 3)
                    d_fr.copy_(flat, non_blocking=True)
                    d_bx = bx.to(self.device, non_blocking=True)
                    d_jr = jr.to(self.device, non_blocking=True) if jr is not None else None
                    ev = torch.cuda.Event()
                    ev.record(self.s_in)
                self._hold.append((flat, bx, jr, ev))  # host sources stay alive until their copy has run
                batch.copied = ev
                self.s_cmp.wait_event(ev)
            else:
                d_fr, d_bx, d_jr = flat, bx, jr
            with (torch.cuda.stream(self.s_cmp) if self.cuda else _null()):
                if self.cuda and self._ev_out[slot] is not None:
                    self.s_cmp.wait_event(self._ev_out[slot])  # the slot's previous block has been downloaded
                # one frame size for the whole batch (the usual case): plain strided copies, no index kernels
                dpos = slice(None) if (len(batch.groups) == 1 and g == B) else pos.to(self.device)
                f = self.extract(d_fr, d_bx, False, None).view(g, T, -1)
                block[dpos, 0] = f.to(self.feat_dtype)
                if self.augment:
                    block[dpos, 3] = torch.flip(f, dims=[1]).to(self.feat_dtype)  # trev == time-reversed orig
                    block[dpos, 1] = self.extract(d_fr, d_bx, False, d_jr).view(g, T, -1).to(self.feat_dtype)
                    block[dpos, 2] = self.extract(d_fr, d_bx, True, None).view(g, T, -1).to(self.feat_dtype)
                    self.stats.frames_computed += 3 * g * T
                else:
                    self.stats.frames_computed += g * T
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(self.s_cmp)
            self._ev_done[slot] = ev
        while len(self._hold) > 8:
            self._hold.pop(0)[3].synchronize()
        return block

    def _d_frames(self, slot: int, nbytes: int) -> torch.Tensor:
        cur = self._dfr[slot]
        if cur is None or cur.numel() < nbytes:
            self._dfr[slot] = cur = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return cur[:nbytes]

    def _d_block(self, slot: int, B: int) -> torch.Tensor:
        cur = self._dblk[slot]
        if cur is None or cur.shape[0] < B:
            self._dblk[slot] = cur = torch.empty(B, self.n_vars, self.T, 2048, dtype=self.feat_dtype,
                                                 device=self.device)
        return cur[:B]

    # ---- the run ------------------------------------------------------------------------------------------------------
    def run(self, feeder_factory: Callable, plan: Sequence[Sequence[int]], my_shards: Sequence[int],
            batch_clips: int) -> PipelineStats:
        """feeder_factory(list of clip-id lists) -> iterable of ClipBatch, one per list, in order."""
        t_all = time.perf_counter()
        self._dfr, self._dblk = [None, None], [None, None]
        self._ev_done, self._ev_out, self._hold = [None, None], [None, None], []
        parts = [(sid, r0, list(plan[sid][r0:r0 + batch_clips])) for sid in my_shards
                 for r0 in range(0, len(plan[sid]), batch_clips)]
        feeder = iter(feeder_factory([p for _, _, p in parts]))
        writer = _EventShardWriter(self.stats)
        cur_sid, buf = None, None
        k = 0
        try:
            for sid, r0, ids in parts:
                t0 = time.perf_counter()
                batch = next(feeder)
                self.stats.feed_wait_s += time.perf_counter() - t0
                assert batch.ids == ids, "feeder returned the wrong clips"
                if sid != cur_sid:
                    cur_sid, buf = sid, self._get_buffers(len(plan[sid]) * self.n_vars, batch)
                B, nv = len(ids), self.n_vars
                slot = k & 1
                k += 1
                block = self._block(batch, slot)
                lo, hi = r0 * nv, (r0 + B) * nv
                if self.cuda:
                    self.s_out.wait_event(self._ev_done[slot])
                    with torch.cuda.stream(self.s_out):
                        buf.feats[lo:hi].copy_(block.view(B * nv, self.T, 2048), non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self.s_out)
                    self._ev_out[slot] = ev
                else:
                    buf.feats[lo:hi].copy_(block.view(B * nv, self.T, 2048))
                    ev = None
                # annotations, vectorised over the batch
                if self.augment:
                    j3, j2, K = augment_annotations_batch(batch.j3, batch.j2, batch.K)
                    buf.j3[lo:hi], buf.j2[lo:hi], buf.K[lo:hi] = j3.flatten(0, 1), j2.flatten(0, 1), K.flatten(0, 1)
                else:
                    buf.j3[lo:hi], buf.j2[lo:hi], buf.K[lo:hi] = batch.j3, batch.j2, batch.K
                for j, i in enumerate(ids):
                    m = self.clip_meta(i)
                    for v in range(nv):
                        buf.metas.append({"subject": m.subject, "action": m.action, "cam": m.cam, "start": m.start,
                                          "end": m.end, "aug": AUG_NAMES[v] if self.augment else "orig",
                                          "box": batch.boxes[j].clone() if not self.augment else None})
                self.stats.clips += B
                if r0 + B >= len(plan[sid]):  # shard complete: hand it to the writer with its last download's event
                    pooled = buf.feats.shape[0] == self._full_rows
                    writer.save({"feats": buf.feats, "joints3d": buf.j3, "joints2d": buf.j2, "K": buf.K,
                                 "meta": buf.metas, "n_vars": nv}, shard_path(self.out_root, sid), ev,
                                (lambda b=buf: self._free.put(b)) if pooled else None,
                                lambda e: self._free.put(e))
                    self.stats.shards += 1
                    self.log(f"rank shard {sid}: {len(plan[sid])} clips queued for writing "
                             f"| {time.perf_counter() - t_all:6.1f}s")
        finally:
            writer.close()
        if self.cuda:
            torch.cuda.current_stream(self.device).synchronize()
        self.stats.total_s = time.perf_counter() - t_all
        return self.stats


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _EventShardWriter:
    """AsyncShardWriter whose thread first waits for the CUDA event of the shard's last D2H copy (so the main thread
    never blocks on the device), then torch.saves in the reference's legacy format and gives the buffer back."""

    def __init__(self, stats: PipelineStats):
        self._q: "queue.Queue" = queue.Queue()
        self._err: Optional[BaseException] = None
        self.stats = stats
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            obj, path, ev, done, on_err = item
            try:
                if self._err is None:
                    if ev is not None:
                        ev.synchronize()
                    t0 = time.perf_counter()
                    torch.save(obj, str(path), _use_new_zipfile_serialization=False)
                    self.stats.write_s += time.perf_counter() - t0
                    self.stats.bytes_written += sum(v.numel() * v.element_size() for v in obj.values()
                                                    if isinstance(v, torch.Tensor))
                if done is not None:
                    done()
            except BaseException as e:  # noqa: BLE001
                self._err = e
                on_err(e)  # unblock a loop waiting for a free buffer

    def save(self, obj, path, ev, done, on_err):
        if self._err:
            raise RuntimeError("shard writer failed") from self._err
        self._q.put((obj, path, ev, done, on_err))

    def close(self):
        self._q.put(None)
        self._t.join()
        if self._err:
            raise RuntimeError("shard writer failed") from self._err
