"""Drop-in entry point: precompute per-clip ResNet-50 features for H36M on B200s.

Same command line and the same on-disk output as the reference's script of the same name
(/root/reference/src/preprocess_resnet_features.py:136-155 flags, :80-131,403-417 shard / index layout), so
`dataset_features.py`, `samplers.py`, `model.py` and `train.py` run unchanged on what this writes.  What changes is
the engine behind `backbone(x)` (:296): libphdfx.so (hand-written sm_100a kernels) instead of cuDNN / DataParallel /
torch.compile, one process per GPU instead of nn.DataParallel (:214-217), pinned async copies instead of the blocking
`.cpu()` (:297), and a write-once shard writer.

    python -u src/preprocess_resnet_features.py --root ROOT --out OUT [--augment] ...           # real data
    torchrun --nproc-per-node 8 src/preprocess_resnet_features.py --root ROOT --out OUT ...     # 8 GPUs
    python -u src/preprocess_resnet_features.py --synthetic 64:224x224 --out OUT --weights random:0   # no dataset

The default writer (--multi-gpu-writer sharded, any number of processes) is the planned one: which clip lands in which
row of which shard is a pure function of (n_clips, shard-size, shuffle-pool, shuffle-seed), so rank r extracts the
clips of shards r, r + world, ... in row order and writes those files itself (phdfx/pipeline.py: batch-level host code,
features downloaded straight into the shard's pinned tensor); rank 0 adds index.pt.  The files hold exactly what the
reference's streaming writer produces.  --multi-gpu-writer gather keeps the streaming writer on rank 0 behind a
feature gather (BASELINE config 4's shape).

Real data needs the user's `dataset.py` (the reference's `Human36MPreprocessedClips`, which owns video decoding and
annotation handling — outside this drop-in) importable, e.g. by running from the reference's src/ directory or with
--dataset-path.  With --seam b (default) the dataset only decodes and computes the person box (phdfx/h36m.py wraps the
user's own methods); crop / resize / normalise and the --augment variants run on the GPU.  --seam a feeds the
dataset's own CPU-preprocessed fp32 clips to the trunk, as the reference does.
Extra flags: --backend {b200,torch}, --synthetic, --weights, --dataset-path, --max-clips, --seam, --feed-threads.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import torch
import torch.nn as nn

_ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(_ROOT / "implementation-phd-lab-vision_b200"))

from phdfx.dist import gather_rows, shard_range  # noqa: E402
from phdfx.shards import (AUG_NAMES, AsyncShardWriter, ClipRecord, ShardWriter, assemble_shard,  # noqa: E402
                          index_from_plan, plan_shards, shard_path)
from phdfx.synthetic import SyntheticH36MClips  # noqa: E402

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def parse_args(argv=None):
    p = argparse.ArgumentParser("B200-native: precompute per-clip ResNet50 features for H36M")
    # --- the reference's 13 flags, same names and defaults (:137-153) ---
    p.add_argument("--root", type=str, default=None, help="H36M preprocessed root")
    p.add_argument("--out", type=str, required=True, help="Output directory for cached features")
    p.add_argument("--seq-len", type=int, default=40)
    p.add_argument("--frame-skip", type=int, default=2)
    p.add_argument("--stride", type=int, default=5)
    p.add_argument("--batch-size", type=int, default=32)
    p.add_argument("--num-workers", type=int, default=8)
    p.add_argument("--subjects", type=int, nargs="+", default=[1, 5, 6, 7, 8, 9, 11])
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--save-fp16", action="store_true", help="Store feats as float16")
    p.add_argument("--augment", action="store_true", help="4 variants per clip: orig, cjitter, hflip, trev")
    p.add_argument("--shard-size", type=int, default=512, help="Number of clips per shard file")
    p.add_argument("--shuffle-pool", type=int, default=8192)
    p.add_argument("--shuffle-seed", type=int, default=123)
    # --- additions ---
    p.add_argument("--backend", choices=["b200", "torch"], default="b200",
                   help="b200: libphdfx.so (no fallback); torch: the reference's eager path (for comparison runs)")
    p.add_argument("--synthetic", type=str, default=None, metavar="N[:HxW[:SIDE[:fast]]]",
                   help="use N synthetic clips of HxW uint8 frames (person box side SIDE, 0 = whole frame) instead of "
                        "--root; ':fast' derives every clip from one random base clip (throughput runs)")
    p.add_argument("--weights", type=str, default="imagenet",
                   help="'imagenet' (torchvision IMAGENET1K_V2, needs network/cache), 'random:SEED', or a state_dict path")
    p.add_argument("--dataset-path", type=str, default=None, help="directory holding the user's dataset.py")
    p.add_argument("--max-clips", type=int, default=None)
    p.add_argument("--jitter-seed", type=int, default=0, help="seed of the colour-jitter variant (synthetic mode)")
    p.add_argument("--multi-gpu-writer", choices=["sharded", "gather"], default="sharded",
                   help="'sharded' = every rank extracts and writes the shards it owns (the shard composition is a pure "
                        "function of the shuffle parameters; no feature gather; also the one-process default), "
                        "'gather' = contiguous clip ranges per rank, features gathered to rank 0, streaming writer")
    p.add_argument("--seam", choices=["a", "b"], default="b",
                   help="real data: b = dataset yields raw uint8 frames + box, GPU does crop/resize/normalise/augment; "
                        "a = dataset's own CPU preprocessing (fp32 clips), as the reference (forces the streaming writer)")
    p.add_argument("--feed-threads", type=int, default=-1,
                   help="--synthetic ...:fast only: host threads that generate clips straight into pinned batch buffers "
                        "(-1 = min(8, cpus / world size), 0 = use the DataLoader workers instead)")
    return p.parse_args(argv)


def build_torch_backbone(weights: str) -> nn.Module:
    """The reference's construction (:207-209)."""
    from torchvision import models

    if weights == "imagenet":
        resnet = models.resnet50(weights=models.ResNet50_Weights.IMAGENET1K_V2)
    elif weights.startswith("random:"):
        torch.manual_seed(int(weights.split(":", 1)[1]))
        resnet = models.resnet50(weights=None)
    else:
        resnet = models.resnet50(weights=None)
        sd = torch.load(weights, map_location="cpu", weights_only=True)
        resnet.load_state_dict(sd.get("state_dict", sd))
    return nn.Sequential(*list(resnet.children())[:-1]).eval()


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            from phdfx.dist import bind_to_gpu_numa

            bind_to_gpu_numa(local)  # before any pinned allocation: host buffers on the GPU's own socket
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return world, rank, local


def real_dataset(args):
    if args.dataset_path:
        sys.path.insert(0, args.dataset_path)
    import torchvision.io as tvio

    if not hasattr(tvio, "VideoReader"):
        # torchvision >= 0.24 dropped VideoReader, which the reference's dataset.py imports by name (:14) and only uses
        # inside a try/except that falls back to torchvision.io.read_video (:323-355): let the import succeed and the
        # fallback do its job
        def _no_video_reader(*a, **k):
            raise RuntimeError("torchvision.io.VideoReader is not available in this torchvision")

        tvio.VideoReader = _no_video_reader
    try:
        explicit = os.path.join(args.dataset_path, "dataset.py") if args.dataset_path else None
        if explicit and os.path.isfile(explicit):
            # --dataset-path names THE file to use: load it even when some other `dataset` module is already imported
            # in this process (an embedding application, a test session), and register it under the name its classes
            # pickle by
            import importlib.util

            spec = importlib.util.spec_from_file_location("dataset", explicit)
            mod = importlib.util.module_from_spec(spec)
            sys.modules["dataset"] = mod
            spec.loader.exec_module(mod)
        from dataset import Human36MPreprocessedClips  # the user's / reference's dataset.py
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(
            "real-data mode needs the reference's dataset.py importable (run from its src/ directory or pass "
            f"--dataset-path); import failed with: {e!r}.  Use --synthetic N to run without a dataset.") from e
    base = Human36MPreprocessedClips(root=args.root, subjects=args.subjects, seq_len=args.seq_len,
                                     frame_skip=args.frame_skip, stride=args.stride,
                                     augment=args.augment and args.seam == "a", max_clips=args.max_clips)
    if args.seam == "a":
        return base
    from phdfx.h36m import U8ClipDataset

    return U8ClipDataset(base)


def parse_synthetic(spec: str, args):
    parts = spec.split(":")
    n = int(parts[0])
    h = w = 224
    side = 0
    if len(parts) > 1:
        h, w = (int(v) for v in parts[1].lower().split("x"))
    if len(parts) > 2:
        side = int(parts[2])
    fast = len(parts) > 3 and parts[3] == "fast"
    if args.max_clips is not None:
        n = min(n, args.max_clips)
    return SyntheticH36MClips(n, seq_len=args.seq_len, height=h, width=w, subjects=tuple(args.subjects), seed=0,
                              box_side=side, fast=fast)


@torch.no_grad()
def main(argv=None):
    args = parse_args(argv)
    world, rank, local = dist_setup()
    is_main = rank == 0
    log = print if is_main else (lambda *a, **k: None)

    if args.backend == "b200":
        if not torch.cuda.is_available():
            raise RuntimeError("--backend b200 needs a CUDA sm_100 device; there is no CPU fallback "
                               "(use --backend torch for the reference's eager path)")
        device = torch.device("cuda", local)
    else:
        # the reference's behaviour (:157-161): fall back to CPU when CUDA is missing
        device = torch.device("cuda", local) if (args.device.startswith("cuda") and torch.cuda.is_available()) \
            else torch.device("cpu")
    n_vars = len(AUG_NAMES) if args.augment else 1
    T = args.seq_len
    out_root = Path(args.out)
    out_root.mkdir(parents=True, exist_ok=True)  # every rank: the ranks write their own shard files
    log(f"Device     : {device}  (world size {world}, backend {args.backend})")
    log(f"Augment    : {args.augment}  ({'4 variants/clip -> ' + ', '.join(AUG_NAMES) if args.augment else 'none'})")
    log(f"Shard size : {args.shard_size} clips  ({args.shard_size * n_vars} variant entries/shard)")

    synthetic = args.synthetic is not None
    if not synthetic and not args.root:
        raise SystemExit("either --root or --synthetic is required")
    if not synthetic and args.multi_gpu_writer == "gather" and args.seam == "b":
        log("--multi-gpu-writer gather streams clips in dataset order, where frame sizes mix within a batch: using --seam a")
        args.seam = "a"
    ds = parse_synthetic(args.synthetic, args) if synthetic else real_dataset(args)
    n_clips = len(ds)
    u8_items = synthetic or args.seam == "b"  # the dataset yields raw uint8 frames + box (Seam B)
    planned = args.multi_gpu_writer == "sharded" and u8_items

    torch_backbone = build_torch_backbone(args.weights)
    if args.backend == "b200":
        import phdfx

        frames_per_call = args.batch_size * T
        backbone = phdfx.B200Backbone(torch_backbone, device=device, max_frames=min(frames_per_call, 2560))
        # host uint8 frames -> host features: pinned, double-buffered H2D / compute / D2H (phdfx/stream.py)
        streamer = phdfx.StreamingExtractor(backbone, batch=min(256, backbone.max_frames))
    else:
        backbone = torch_backbone.to(device)
        streamer = None
    feat_dtype = torch.float16 if args.save_fp16 else torch.float32

    def run_normalised(v_video: torch.Tensor) -> torch.Tensor:
        """(Bv,T,3,224,224) fp32 normalised -> (Bv,T,2048) on the device — the reference's lines :288-296."""
        v_video = v_video.to(device, non_blocking=True)
        Bv, Tt, C, H, W = v_video.shape
        x = v_video.view(Bv * Tt, C, H, W).contiguous()
        if args.backend == "torch":
            with torch.autocast(device_type=device.type, dtype=torch.bfloat16, enabled=device.type == "cuda"):
                return backbone(x).flatten(1).view(Bv, Tt, -1).float()
        return backbone(x).flatten(1).view(Bv, Tt, -1)

    extract_fn = make_extract(args, backbone, device, T)

    def run_u8(frames: torch.Tensor, boxes: torch.Tensor, flip: bool, to_host: bool = False) -> torch.Tensor:
        """Seam B (streaming-writer path): (Bv,T,H,W,3) uint8 + per-clip boxes -> (Bv,T,2048) on the device (to_host: in
        pinned host memory, through the streaming extractor)."""
        Bv, Tt, H, W, _ = frames.shape
        bx = boxes.to(torch.int32).repeat_interleave(Tt, dim=0)
        if to_host and args.backend == "b200":  # overlapped copies, features land in pinned host memory
            return streamer(frames.view(Bv * Tt, H, W, 3), bx, flip_w=flip).view(Bv, Tt, -1)
        fr = frames.view(Bv * Tt, H, W, 3).to(device, non_blocking=True)
        return extract_fn(fr, bx.to(device, non_blocking=True), flip, None).view(Bv, Tt, -1)

    def jitter_variant(frames: torch.Tensor, boxes: torch.Tensor, clip_ids) -> torch.Tensor:
        """Colour-jitter variant — the reference's recipe (src/dataset.py:188-198: ColorJitter on the resized [0,1]
        clip, one draw per clip, then Normalize) with the per-clip draws of jitter_rows; b200: inside K1
        (phdfx_extract_u8_jitter), torch: torchvision's own ops."""
        Bv, Tt, H, W, _ = frames.shape
        jrows = jitter_rows(args.jitter_seed, clip_ids).repeat_interleave(Tt, dim=0).to(device, non_blocking=True)
        fr = frames.view(Bv * Tt, H, W, 3).to(device, non_blocking=True)
        bx = boxes.to(torch.int32).repeat_interleave(Tt, dim=0).to(device, non_blocking=True)
        return extract_fn(fr, bx, False, jrows).view(Bv, Tt, -1)

    def extract_clips(ids, items, to_host: bool = False):
        """Features + annotations of the clips `ids` (dataset numbers; `items` = what the dataset returned for them,
        synthetic clips already collated): (feats [len(ids), n_vars, T, 2048] in feat_dtype — on the device, or on the
        host when to_host —, [(joints3d[v], joints2d[v], K[v], box)])."""
        feats = torch.empty(0, n_vars, T, 2048, device="cpu" if to_host else device)
        small = []  # per clip: (joints3d[v], joints2d[v], K[v], box)
        if not ids:
            return feats.to(feat_dtype), small
        if u8_items:
            frames, boxes, items = items  # (Bv,T,H,W,3) uint8, (Bv,4), per-clip tuples without the frames
            host = to_host and args.backend == "b200"
            f_orig = run_u8(frames, boxes, False, to_host=host)
            if args.augment:
                f_jit = jitter_variant(frames, boxes, ids)
                f_flip = run_u8(frames, boxes, True, to_host=host)
                if host:
                    f_jit = f_jit.cpu()
                f_trev = torch.flip(f_orig, dims=[1])  # frames are independent: exact
                feats = torch.stack([f_orig, f_jit, f_flip, f_trev], dim=1)
            else:
                feats = f_orig.unsqueeze(1)
            for it in items:
                j3, j2, K, box = it
                if args.augment:
                    small.append(_augment_annotations(j3, j2, K))
                else:
                    small.append(([j3], [j2], [K], box))
        else:
            # Seam A: clips were stacked per variant in the loader's workers (_collate_seam_a) and pinned by its
            # pin-memory thread, so the copies below are asynchronous (the reference collates in its workers too, :59-69)
            vids, small = items
            if args.augment:  # variants as src/dataset.py:411-426 builds them
                f = [run_normalised(vids[0]), run_normalised(vids[1]), run_normalised(vids[2])]
                f.append(torch.flip(f[0], dims=[1]))  # trev == reversed orig (dataset.py:201-207)
                feats = torch.stack(f, dim=1)
            else:
                feats = run_normalised(vids[0]).unsqueeze(1)
        feats = feats.to(feat_dtype)
        return (feats.cpu() if to_host else feats), small

    def clip_batches(batches):
        """(ids, items) for every id list in `batches`, in order.  With --num-workers > 0 one persistent DataLoader
        decodes / generates the clips of batch i+1.. in worker processes (pinned) while batch i is on the GPU — the
        reference's worker pool (:195-204), driven by an explicit batch list instead of a sequential sampler."""
        live = [b for b in batches if b]
        if args.num_workers > 0 and live:
            from torch.utils.data import DataLoader

            loader = iter(DataLoader(ds, batch_sampler=live, num_workers=min(args.num_workers, len(live)),
                                     collate_fn=_collate_synthetic if u8_items else _collate_seam_a,
                                     pin_memory=torch.cuda.is_available(), prefetch_factor=2))
        else:
            loader = ((_collate_synthetic if u8_items else _collate_seam_a)([ds[i] for i in b]) for b in live)
        for b in batches:
            yield b, (next(loader) if b else None)

    def clip_record(i, host_feats, small_i):
        """ClipRecord of dataset clip i: host_feats [n_vars, T, 2048], small_i as returned by extract_clips."""
        clip = ds.index[i]
        j3s, j2s, Ks, box = small_i
        metas = [{"subject": clip.subject, "action": clip.action, "cam": clip.cam, "start": clip.start,
                  "end": clip.end, "aug": AUG_NAMES[v] if args.augment else "orig",
                  "box": box if not args.augment else None} for v in range(n_vars)]
        return ClipRecord([host_feats[v] for v in range(n_vars)], j3s, j2s, Ks, metas)

    B = args.batch_size
    t_all = time.time()
    if planned:
        # ---- planned writer: every rank extracts the shards it owns, in row order, and writes them itself -----------
        from phdfx.pipeline import ClipFeeder, PlannedShardRun

        plan = plan_shards(n_clips, args.shard_size, args.shuffle_pool, args.shuffle_seed)
        mine = list(range(rank, len(plan), world))
        if is_main:  # shard files of an earlier run with a different plan would sit next to the new index
            for old in sorted(out_root.glob("shard_*.pt")):
                try:
                    if int(old.stem.split("_")[1]) >= len(plan):
                        log(f"removing stale {old.name} (this run writes {len(plan)} shard(s))")
                        old.unlink()
                except (ValueError, OSError):
                    pass
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        run = PlannedShardRun(extract=make_extract(args, backbone, device, T), device=device, out_root=out_root,
                              n_vars=n_vars, seq_len=T, feat_dtype=feat_dtype, augment=args.augment,
                              jitter_rows=lambda ids: jitter_rows(args.jitter_seed, ids),
                              clip_meta=lambda i: ds.index[i], log=log)
        threads = args.feed_threads if args.feed_threads >= 0 else max(2, min(8, (os.cpu_count() or 8) // world))
        threads = threads if (synthetic and getattr(ds, "fast", False)) else 0
        st = run.run(lambda batches: ClipFeeder(ds, batches, workers=args.num_workers,
                                                pin=device.type == "cuda", threads=threads),
                     plan, mine, B)
        if world > 1:
            dist.barrier()
        if is_main:
            index = index_from_plan(plan, lambda i: ds.index[i], n_vars, args.seq_len, args.frame_skip,
                                    args.save_fp16, args.augment, args.shuffle_seed, args.shuffle_pool)
            torch.save(index, out_root / "index.pt")
            total = time.time() - t_all
            log("-" * 60)
            log(f"Done: {n_clips} clips x {n_vars} variant(s) packed into {len(plan)} shard(s) by {world} ranks")
            log(f"Total time {total:.1f}s | {n_clips / max(total, 1e-9):.1f} clips/s "
                f"({n_clips * n_vars * T / max(total, 1e-9):.0f} frames/s)")
            log(f"rank 0 loop: {st.clips} clips, {st.frames_computed} frames through the trunk in {st.total_s:.2f}s "
                f"({st.frames_computed / max(st.total_s, 1e-9):.0f} computed frames/s, "
                f"{st.clips * n_vars * T / max(st.total_s, 1e-9):.0f} written frames/s) | waited {st.feed_wait_s:.2f}s "
                f"for frames, {st.writer_wait_s:.2f}s for the writer | writer: {st.bytes_written / 1e6:.0f} MB in "
                f"{st.write_s:.2f}s ({st.bytes_written / 1e6 / max(st.write_s, 1e-9):.0f} MB/s, overlapped)")
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    writer = ShardWriter(out_root, n_vars, args.shard_size, args.shuffle_pool, args.shuffle_seed) if is_main else None

    # clips are dealt out in global batches of (world * batch_size): rank r takes the r-th contiguous slice, so
    # rank 0 can append clips to the shuffle pool in the reference's order after every gather
    t_last = t_all
    done = 0
    starts = list(range(0, n_clips, B * world))
    my_ids = []
    for g0 in starts:
        lo, hi = shard_range(min(n_clips, g0 + B * world) - g0, rank, world)
        my_ids.append(list(range(g0 + lo, g0 + hi)))
    fetched = clip_batches(my_ids)
    for g0 in starts:
        g1 = min(n_clips, g0 + B * world)
        ids, items = next(fetched)
        feats, small = extract_clips(ids, items, to_host=world == 1)
        all_feats = gather_rows(feats.contiguous(), g1 - g0, dst=0)
        if world > 1:
            import torch.distributed as dist

            gathered = [None] * world if is_main else None
            dist.gather_object(small, gathered, dst=0)
            small_all = [s for part in gathered for s in part] if is_main else None
        else:
            small_all = small
        if is_main:
            host = all_feats.cpu()
            for k in range(g1 - g0):
                writer.add(clip_record(g0 + k, host[k], small_all[k]))
            done = g1
            if done % 200 < B * world or done == n_clips:
                dt = time.time() - t_last
                t_last = time.time()
                log(f"[{100 * done / n_clips:5.1f}%] {done:6d}/{n_clips} clips | shard {writer.shard_id} "
                    f"(pool: {len(writer.pool)} clips, carry: {len(writer.carry)} clips) | {dt:5.2f}s")
    if is_main:
        log("\nWaiting for all shards to be written to disk...")
        writer.finish(seq_len=args.seq_len, frame_skip=args.frame_skip, save_fp16=args.save_fp16,
                      augment=args.augment)
        total = time.time() - t_all
        log("-" * 60)
        log(f"Done: {n_clips} clips x {n_vars} variant(s) packed into {writer.shard_id} shard(s)")
        log(f"Total time {total:.1f}s | {n_clips / max(total, 1e-9):.1f} clips/s "
            f"({n_clips * n_vars * T / max(total, 1e-9):.0f} frames/s)")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


def jitter_rows(jitter_seed: int, clip_ids) -> torch.Tensor:
    """One row of colour-jitter parameters per clip (fp32 [B,12], phdfx.jitter_params layout): the draw torchvision's own
    ColorJitter(0.3, 0.3, 0.2, 0.05).make_params makes (the reference's recipe, src/dataset.py:188-198, one draw per
    clip) under a per-clip seed, so a run is reproducible and independent of how clips are batched or sharded."""
    from torchvision.transforms import v2 as T2

    from phdfx.backbone import jitter_params

    jit = T2.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.2, hue=0.05)
    rows = []
    for i in clip_ids:
        torch.manual_seed(jitter_seed * 1_000_003 + int(i))
        prm = jit.make_params([])
        rows.append(jitter_params(prm["fn_idx"], prm["brightness_factor"], prm["contrast_factor"],
                                  prm["saturation_factor"], prm["hue_factor"]))
    return torch.stack(rows)


def make_extract(args, backbone, device, T):
    """extract(frames [N,H,W,3] uint8, boxes [N,4] int32, flip, jitter [N,12] | None) -> fp32 [N,2048], all on `device`
    (phdfx/pipeline.py).  b200: one call into libphdfx (K1 + trunk).  torch: the reference's own front end clip by clip
    (src/dataset.py:141-152, :166, :188-198, :242-245) and its eager trunk — the comparison arm."""
    if args.backend == "b200":
        return lambda frames, boxes, flip, jitter: backbone.extract_u8(frames, boxes, flip_w=flip, jitter=jitter)

    import torchvision.transforms.functional as TF
    from torchvision.transforms.v2 import functional as F2

    mean = torch.tensor(IMAGENET_MEAN, device=device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=device).view(1, 3, 1, 1)
    ops = (F2.adjust_brightness, F2.adjust_contrast, F2.adjust_saturation, F2.adjust_hue)

    def extract(frames, boxes, flip, jitter):
        vids = []
        for c0 in range(0, frames.shape[0], T):
            top, left, hh, ww = (int(v) for v in boxes[c0])
            v = frames[c0:c0 + T].permute(0, 3, 1, 2)[:, :, top:top + hh, left:left + ww]
            v = TF.resize(v, [224, 224], antialias=False).to(torch.float32) / 255.0
            if jitter is not None:  # ColorJitter.transform with the given draw (transforms/v2/_color.py:156-173)
                row = jitter[c0].tolist()
                factor = {0: row[4], 1: row[5], 2: row[7], 3: row[9]}
                for k in range(4):
                    v = ops[int(row[k])](v, factor[int(row[k])])
            if flip:
                v = torch.flip(v, dims=[-1])
            vids.append((v - mean) / std)
        x = torch.cat(vids).to(device)
        with torch.autocast(device_type=device.type, dtype=torch.bfloat16, enabled=device.type == "cuda"):
            return backbone(x).flatten(1).float()

    return extract


def _augment_annotations(j3, j2, K):
    """Annotation side of the four variants (src/dataset.py:158-207): hflip mirrors x and swaps left/right joints."""
    flip_pairs = [(1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13)]  # src/dataset.py:39-46
    j2f, j3f, Kf = j2.clone(), j3.clone(), K.clone()
    j2f[..., 0] = 224 - j2f[..., 0]
    j3f[..., 0] = -j3f[..., 0]
    for l, r in flip_pairs:
        j2f[:, [l, r]] = j2f[:, [r, l]]
        j3f[:, [l, r]] = j3f[:, [r, l]]
    Kf[0, 2] = 224 - Kf[0, 2]
    return ([j3, j3, j3f, torch.flip(j3, dims=[0])], [j2, j2, j2f, torch.flip(j2, dims=[0])], [K, K, Kf, K], None)


def _collate_seam_a(batch):
    """Seam A clips (the reference dataset's own items, src/dataset.py:403-437), collated in the worker: one stacked
    fp32 tensor per computed variant (so the loader's pin-memory thread page-locks whole batches) + per-clip
    annotations.  The time-reversed variant's pixels are dropped here: its features are the reversed `orig` features."""
    if isinstance(batch[0], list):  # --augment: 4 x (video, j3d, j2d, K)
        vids = [torch.stack([it[v][0] for it in batch]) for v in range(3)]
        small = [([it[v][1] for v in range(4)], [it[v][2] for v in range(4)], [it[v][3] for v in range(4)], None)
                 for it in batch]
        return vids, small
    return [torch.stack([it[0] for it in batch])], [([it[1]], [it[2]], [it[3]], it[4]) for it in batch]


def _collate_synthetic(batch):
    """Synthetic clips: stack the uint8 frames and boxes in the worker, keep the small annotations per clip."""
    return (torch.stack([it[0] for it in batch]), torch.stack([it[4] for it in batch]),
            [(it[1], it[2], it[3], it[4]) for it in batch])


if __name__ == "__main__":
    main()
