"""bench.py — ResNet-50 frame-feature extraction throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): one synthetic H36M-shaped sequence of 2000 uint8 frames 224x224x3, batch 256.
A step = one pass of the hot path (K1 preprocess -> fused stem+maxpool -> 52 bottleneck convs, the last with the
average pool fused; layer1/layer2 blocks run as fused conv2 -> conv3 [-> conv1] chains) over one batch of
256 frames taken cyclically from the sequence.  The sequence (301 MB) is resident in HBM and larger than the 126 MB
L2, and consecutive steps read different batches, so inputs come from HBM every step.

One JSON line on stdout (rank 0):
  value         frames/s over all ranks, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e           the same metric through the public host-buffer API (phdfx.StreamingExtractor): pinned host uint8
                frames -> H2D -> features -> D2H, copies inside the timed region
  roofline      trunk (every launch of a step except K1) FLOP/s against the measured dense-bf16 peak; 2*MAC
                convention: 8.174 GFLOP per frame (4 087 136 256 MAC).  `frac` is against the
                BURST peak (the timed region lasts tens of milliseconds); `sustained` is a separate >= 2 s leg with its
                own clock record against the sustained peak.  `traffic` comes from the committed ncu capture whose
                csrc hash matches the running sources (`traffic_stale` says when it does not).
  gpu_baseline  the reference's own GPU path on this B200 (torchvision ResNet-50 minus fc, eager cuDNN under bf16
                autocast, src/preprocess_resnet_features.py:164-167,207-209,288-297), timed in a child process so the
                product process never loads cuDNN
  cpu_baseline  a port of the reference's CPU path (same torchvision calls, fp32 eager) on this box's host cores
  config4       (N > 1) BASELINE config 4: 200 000 frames frame-range-sharded over the ranks, final NCCL gather of the
                (200000, 2048) fp32 features to rank 0
--impl reference times the CPU path alone with the same JSON contract.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

FLOP_PER_FRAME = 2 * 4_087_136_256  # all 53 convs, 2*MAC (same convention as MEASURED_PEAKS.json's 2*N^3)
SEQ_FRAMES = 2000
BATCH = 256
IMG = 224
METRIC = "resnet50_feature_frames_per_s"
UNIT = "frames/s"
CONFIG4_FRAMES = 200_000
SUSTAINED_SECONDS = 2.5


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region through NVML (pynvml, ~1 kHz capable; the
    nvidia-smi CLI takes longer per query than a whole timed region lasts).  Falls back to nvidia-smi."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []  # (sm_mhz, reasons_bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                if nv is not None:
                    mhz = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    except Exception:  # noqa: BLE001
                        reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    self.samples.append((mhz, reasons))
                    self._stop.wait(0.004)
                else:
                    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                          "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append((float(out[0]), 0))
                    self.sm_max = float(out[1])
                    self._stop.wait(0.05)
            except Exception:  # noqa: BLE001
                self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        # NVML clocks-event-reason bits (nvml.h): 0x4 sw_power_cap, 0x8 hw_slowdown, 0x20 sw_thermal_slowdown,
        # 0x40 hw_thermal_slowdown, 0x80 hw_power_brake_slowdown
        names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown"}
        sm = [s[0] for s in self.samples]
        mask = 0
        for s in self.samples:
            mask |= s[1]
        reasons = sorted(n for b, n in names.items() if mask & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(sm), "sm_mhz_min": float(min(sm)) if sm else None,
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------------------
# comparison arms (never on the product path)
# ---------------------------------------------------------------------------------------------------------------
def reference_cpu_steps(steps: int, warmup: int, frames_per_step: int):
    """A PORT of the reference's CPU implementation of the path, on all host cores: the same torchvision calls in the
    same order — crop/resize/normalise as src/dataset.py:141-152,242-245 (F.resize on uint8, /255, Normalize) and
    backbone(x).flatten(1) with the trunk built exactly as src/preprocess_resnet_features.py:207-209 (fp32 eager; the
    reference disables autocast / compile / DataParallel on CPU, :157-161,220,239-241).  The reference's own module
    cannot be imported on the GPU box (/root/reference does not travel, and it has no installable package), hence
    kind = "port"."""
    import torchvision.transforms.functional as TF

    from phdfx.synthetic import seeded_backbone, seeded_frames

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    backbone = seeded_backbone()
    frames = torch.from_numpy(seeded_frames(frames_per_step, IMG, IMG, 2))
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)

    def step():
        with torch.no_grad():
            v = frames.permute(0, 3, 1, 2)
            v = TF.resize(v, [IMG, IMG], antialias=False).to(torch.float32) / 255.0
            x = (v - mean) / std
            return backbone(x).flatten(1)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"fps": frames_per_step * steps / dt, "ms_per_step": dt / steps * 1e3, "cores": cores,
            "sample": f"{steps} steps x {frames_per_step} frames 224x224 (crop/resize/normalise + trunk), fp32 eager "
                      f"torchvision, {warmup} warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps_frames = 32
    r = reference_cpu_steps(max(1, args.steps), max(1, args.warmup), fps_frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic H36M sequence, 224x224 uint8 frames, ResNet-50 2048-d features; "
                               f"reference CPU path, bounded sample of {fps_frames} frames per step"},
        "cpu_baseline": {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"]},
        "e2e": {"value": r["fps"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """The reference's GPU path ("second baseline"), one JSON line.  Run as a child process of the
    b200 arm (`--impl reference-gpu`): torchvision ResNet-50 minus fc in eval mode, eager cuDNN under
    torch.autocast(bf16) with cudnn.benchmark and TF32 allowed (src/preprocess_resnet_features.py:164-167), fp32 NCHW
    normalised input resident on the device, `backbone(x).flatten(1)` (:207-209, :288-297), batch 256."""
    from phdfx.synthetic import seeded_backbone

    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    backbone = seeded_backbone().to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(2)
    xs = [torch.randn(BATCH, 3, IMG, IMG, device=dev, generator=g) for _ in range(2)]

    def fwd(x):
        with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            return backbone(x).flatten(1).float()

    for i in range(5):  # cudnn.benchmark picks its algorithms here
        fwd(xs[i & 1])
    torch.cuda.synchronize(dev)
    reps = max(5, min(args.steps, 20))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fwd(xs[i & 1])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"impl": "reference-gpu", "value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                      "steps": reps, "batch": BATCH, "dtype": "bf16 autocast (fp32 in/out)",
                      "kind": "the reference's GPU path: torchvision ResNet-50 minus fc, eager cuDNN "
                              f"{torch.backends.cudnn.version()} under torch.autocast(bf16), cudnn.benchmark, input "
                              "resident in HBM (src/preprocess_resnet_features.py:164-167,207-209,288-297)"}),
          flush=True)


def gpu_baseline_child(local: int, steps: int):
    env = dict(os.environ)
    env["LOCAL_RANK"] = str(local)
    for k in ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference-gpu", "--steps", str(steps)],
                             capture_output=True, text=True, timeout=240, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                d.pop("impl", None)
                return d
        return {"unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:200]}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}


def committed_traffic(sha: str):
    """DRAM traffic of the trunk's launches from the committed `ncu --set full` capture of one step (bench.py cannot
    read dram__bytes counters itself).  The newest capture wins; `stale` = its csrc hash is not the running one."""
    import re

    def _ver(pth):  # profiles/rNN/trunk_traffic_vMM.json -> (NN, MM)
        m = re.search(r"r(\d+)[/\\]trunk_traffic_v(\d+)", str(pth))
        return (int(m.group(1)), int(m.group(2))) if m else (0, 0)

    traffic = None
    for cand in sorted((ROOT / "profiles").glob("r*/trunk_traffic_*.json"), key=_ver):
        try:
            traffic = json.loads(cand.read_text())
            traffic["file"] = str(cand.relative_to(ROOT))
        except Exception:  # noqa: BLE001
            pass
    if traffic is not None:
        traffic["stale"] = traffic.get("csrc_sha") != sha
    return traffic


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 4 (runs when WORLD_SIZE > 1)
# ---------------------------------------------------------------------------------------------------------------
def config4_leg(eng, dev, rank: int, world: int, total: int = CONFIG4_FRAMES):
    """200 000 frames (224x224 uint8, generated per rank on the device, seed 4 + rank), contiguous frame ranges over
    the ranks, batch 256 (one CUDA-graph replay per full batch over a fixed input slot), final NCCL gather of the
    (200000, 2048) fp32 features to rank 0.  Time = max over ranks of extraction + gather (CUDA events)."""
    import torch.distributed as dist

    from phdfx.dist import shard_range

    lo, hi = shard_range(total, rank, world)
    n = hi - lo
    per = (total + world - 1) // world
    g = torch.Generator(device=dev).manual_seed(4 + rank)
    frames = torch.empty(n, IMG, IMG, 3, dtype=torch.uint8, device=dev)  # 150.5 KB per frame: 15 GB at 100k frames
    for c0 in range(0, n, 2048):
        c1 = min(n, c0 + 2048)
        frames[c0:c1] = torch.randint(0, 256, (c1 - c0, IMG, IMG, 3), dtype=torch.uint8, device=dev, generator=g)
    feats = torch.zeros(per, 2048, dtype=torch.float32, device=dev)  # padded to the common range length
    dst = torch.empty(world * per, 2048, dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None

    def gather():
        if world > 1:
            dist.gather(feats, list(dst.chunk(world)) if rank == 0 else None, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    slot = torch.empty(BATCH, IMG, IMG, 3, dtype=torch.uint8, device=dev)
    out_slot = torch.empty(BATCH, 2048, dtype=torch.float32, device=dev)
    slot.copy_(frames[:BATCH])
    graph = eng.capture_extract(slot, None, out=out_slot)
    graph.replay()
    gather()
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    launches = 0
    with ClockSampler(dev.index) as clocks:
        barrier()
        e0.record()
        for b0 in range(0, n, BATCH):
            b1 = min(n, b0 + BATCH)
            if b1 - b0 == BATCH:
                slot.copy_(frames[b0:b1])
                graph.replay()
                feats[b0:b1].copy_(out_slot)
                launches += graph.launches
            else:
                eng.extract_u8(frames[b0:b1], None, out=feats[b0:b1])
                launches += eng.launches
        e1.record()
        gather()
        e2.record()
        barrier()
    t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1), e1.elapsed_time(e2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_extract, ms_gather = (float(v) for v in t.tolist())
    ok = True
    if rank == 0:
        ok = bool(torch.isfinite(dst if dst is not None else feats).all().item())
    del frames, feats, dst
    torch.cuda.empty_cache()
    return {"workload": "configs[3]: 200 000 frames 224x224 uint8 generated on the devices, contiguous frame ranges, "
                        "batch 256, final NCCL gather of (200000, 2048) fp32 to rank 0",
            "frames": total, "n_gpus": world, "frames_per_s": total / (ms_total / 1e3),
            "frames_per_s_per_gpu_extract_only": n / (ms_extract / 1e3), "ms_total_max_over_ranks": ms_total,
            "ms_extract_max_over_ranks": ms_extract, "ms_gather_max_over_ranks": ms_gather,
            "gather_bytes": (world - 1) * per * 2048 * 4 if world > 1 else 0, "gpu_launches_rank0": launches,
            "scaling": "strong", "finite": ok, "clocks_rank0": clocks.summary()}


# ---------------------------------------------------------------------------------------------------------------
# the product arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist

    import phdfx
    from phdfx.synthetic import csrc_sha, seeded_backbone

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from phdfx.dist import bind_to_gpu_numa

    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None  # pinned host buffers on the GPU's own socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    K, Wm = args.steps, args.warmup
    backbone = seeded_backbone()
    eng = phdfx.B200Backbone(backbone, device=local, max_frames=BATCH)
    # the sequence: generated on the device (config 2), seed 2 + rank; each rank owns its own frame range
    g = torch.Generator(device=dev).manual_seed(2 + rank)
    seq = torch.randint(0, 256, (SEQ_FRAMES, IMG, IMG, 3), dtype=torch.uint8, device=dev, generator=g)
    n_batches = SEQ_FRAMES // BATCH  # 7 full batches used cyclically (1792 frames = 270 MB > L2)
    feats_all = torch.empty(K, BATCH, 2048, dtype=torch.float32, device=dev)
    scratch = torch.empty(BATCH, 2048, dtype=torch.float32, device=dev)

    def step(i, out):
        b = i % n_batches
        eng.extract_u8(seq[b * BATCH:(b + 1) * BATCH], None, out=out)

    # one CUDA graph per timed step (input slice -> that step's feature rows), captured up front
    graphs = [eng.capture_extract(seq[((Wm + i) % n_batches) * BATCH:((Wm + i) % n_batches + 1) * BATCH], None,
                                  out=feats_all[i]) for i in range(K)]

    def gather_all():
        if world > 1:
            dst = torch.empty(world * K * BATCH, 2048, dtype=torch.float32, device=dev) if rank == 0 else None
            dist.gather(feats_all.view(-1, 2048), list(dst.chunk(world)) if rank == 0 else None, dst=0)

    for i in range(Wm):
        step(i, scratch)
    for gph in graphs:  # untimed: the first replay of a graph also uploads it to the device
        gph.replay()
    gather_all()  # warm-up: NCCL sets up its peer connections lazily on the first point-to-point operation
    barrier()

    # ---- device-resident throughput (value) ----------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for i in range(K):
            graphs[i].replay()
            launches += graphs[i].launches
        gather_all()  # the only collective of the path: final feature gather to rank 0 over NVLink
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    fps = world * K * BATCH / (ms_max / 1e3)

    # ---- trunk-only timing for the roofline (same steps, events around phdfx_forward only) -------------------
    x4 = eng.preprocess_u8(seq[:BATCH], None)
    for _ in range(2):
        eng.forward_nhwc4p(x4)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(K, 5)
    x4s = [eng.preprocess_u8(seq[b * BATCH:(b + 1) * BATCH], None) for b in range(4)]  # 4 x 106 MB inputs > L2
    trunk_graphs = [eng.capture(lambda x=x: eng.forward_nhwc4p(x)) for x in x4s]  # same launch path as `value`
    for gph in trunk_graphs:
        gph.replay()
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        trunk_graphs[i % 4].replay()
    e1.record()
    torch.cuda.synchronize(dev)
    trunk_ms = e0.elapsed_time(e1) / reps
    pk = peaks()
    achieved_tf = BATCH * FLOP_PER_FRAME / (trunk_ms / 1e3) / 1e12

    # ---- sustained legs (>= SUSTAINED_SECONDS each, own clock record): the trunk alone, then whole steps -------
    def sustained(replay, ms_guess):
        n_rep = max(16, int(SUSTAINED_SECONDS * 1e3 / ms_guess) + 1)
        with ClockSampler(local) as ck:
            torch.cuda.synchronize(dev)
            e0.record()
            for i in range(n_rep):
                replay(i)
            e1.record()
            torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n_rep, n_rep, ck.summary()

    sus_trunk_ms, sus_trunk_n, sus_trunk_ck = sustained(lambda i: trunk_graphs[i % 4].replay(), trunk_ms)
    sus_step_ms, sus_step_n, sus_step_ck = sustained(lambda i: graphs[i % K].replay(), ms_max / K)
    sus_tf = BATCH * FLOP_PER_FRAME / (sus_trunk_ms / 1e3) / 1e12
    sha = csrc_sha()
    traffic = committed_traffic(sha)

    # ---- K1 alone (HBM-bound) ----------------------------------------------------------------------------------
    k1_out = x4s[0]
    for i in range(2):
        eng.preprocess_u8(seq[i * BATCH:(i + 1) * BATCH], None, out=k1_out)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        eng.preprocess_u8(seq[(i % n_batches) * BATCH:(i % n_batches + 1) * BATCH], None, out=k1_out)
    e1.record()
    torch.cuda.synchronize(dev)
    k1_ms = e0.elapsed_time(e1) / reps
    k1_bytes = BATCH * (IMG * IMG * 3 + IMG * 232 * 4 * 2)
    k1_gbs = k1_bytes / (k1_ms / 1e3) / 1e9
    # K1 with the colour-jitter variant (two launches: grey-level row sums, then jitter + normalise): reads the
    # frames twice, writes the same output
    jrow = phdfx.jitter_params((1, 3, 0, 2), 1.18, 0.77, 1.2, -0.02)
    jrows = jrow.unsqueeze(0).repeat(BATCH, 1).to(dev)
    for i in range(2):
        eng.preprocess_u8(seq[i * BATCH:(i + 1) * BATCH], None, out=k1_out, jitter=jrows)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(reps):
        eng.preprocess_u8(seq[(i % n_batches) * BATCH:(i % n_batches + 1) * BATCH], None, out=k1_out, jitter=jrows)
    e1.record()
    torch.cuda.synchronize(dev)
    k1j_ms = e0.elapsed_time(e1) / reps
    k1j_bytes = BATCH * (2 * IMG * IMG * 3 + IMG * 232 * 4 * 2)

    # ---- end to end through the host-buffer API ----------------------------------------------------------------
    # the host-side sequence: K batches (at most 32 = 1.2 GB pinned) taken cyclically from the same frames, so ONE call
    # streams the whole timed region (upload of batch i+1 / download of batch i-1 overlap the trunk of batch i, and
    # the pipeline fills and drains once, as it does for a real sequence)
    nb_e2e = max(2, min(K, 32))
    host = torch.empty(nb_e2e * BATCH, IMG, IMG, 3, dtype=torch.uint8).pin_memory()
    for b in range(nb_e2e):
        host[b * BATCH:(b + 1) * BATCH].copy_(seq[(b % n_batches) * BATCH:(b % n_batches + 1) * BATCH])
    torch.cuda.synchronize(dev)
    se = phdfx.StreamingExtractor(eng, batch=BATCH)
    warm_out = torch.empty(2 * BATCH, 2048, dtype=torch.float32).pin_memory()
    se(host[:2 * BATCH], None, out=warm_out)  # warm-up: allocates staging, captures the graph of both slots
    barrier()
    frames_k = host
    reps_e2e = max(1, (K + nb_e2e - 1) // nb_e2e)
    out_k = torch.empty(frames_k.shape[0], 2048, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(reps_e2e):
        se(frames_k, None, out=out_k)
        h2d += se.h2d_bytes
        d2h += se.d2h_bytes
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    n_e2e_frames = reps_e2e * frames_k.shape[0]
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_fps = world * n_e2e_frames / float(te.item())
    steps_e2e = n_e2e_frames / BATCH
    del host, out_k, warm_out, se

    # ---- BASELINE config 4 under torchrun ------------------------------------------------------------------------
    cfg4 = None
    if world > 1 and not args.no_config4:
        cfg4 = config4_leg(eng, dev, rank, world)

    n_chain = sum(1 for i in range(len(eng.plan.layers)) if eng.chain_span(i) > 0)
    n_trunk = trunk_graphs[0].launches  # phdfx_forward on an NHWC4p tensor: the stem without K1 inside
    sched, sched_flags = eng.get_schedule()
    if rank == 0:
        cpu = gpu_ref = None
        if world == 1 and not args.no_cpu_baseline:
            r = reference_cpu_steps(4, 1, 32)
            cpu = {"value": r["fps"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if world == 1 and not args.no_gpu_baseline:
            del graphs, trunk_graphs, x4s, feats_all, seq
            torch.cuda.empty_cache()
            gpu_ref = gpu_baseline_child(local, K)
        line = {
            "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1]: single synthetic H36M sequence, 2000 frames 224x224 uint8, batch 256; "
                                   "random-init ResNet-50 (seeded) with seeded BN stats",
                       "batch": BATCH, "frames_per_rank_per_step": BATCH,
                       "launch": f"each step is one CUDA-graph replay of its {launches // K} kernel launches (PDL edges inside)",
                       "schedule": {"stages_first_layer_wave_frames": sched, "flags": sched_flags},
                       "l2": "inputs larger than L2: steps cycle over 7 batches of a 301 MB HBM-resident sequence",
                       "parallelism": f"frame-range sharding x{world}, no collective on the math path"
                                      + (", final NCCL gather of features inside the timed region" if world > 1 else "")},
            "e2e": {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": int(h2d / steps_e2e),
                    "d2h_bytes_per_step": int(d2h / steps_e2e),
                    "api": "phdfx.StreamingExtractor (pinned host uint8 -> features in pinned host fp32)",
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": pk["bf16_burst"], "unit": "TFLOP/s",
                         "frac": achieved_tf / pk["bf16_burst"],
                         "regime": f"burst: {reps} graph replays of the trunk = {trunk_ms * reps:.0f} ms, measured against "
                                   "the burst peak (MEASURED_PEAKS.json bf16_tflops)",
                         "traffic": traffic["traffic_bytes_per_step"] if traffic else None,
                         "traffic_source": (traffic["file"] + " (ncu dram__bytes_read+write summed over the trunk's "
                                            "launches of one step)") if traffic else None,
                         "traffic_csrc_sha": traffic.get("csrc_sha") if traffic else None,
                         "traffic_stale": traffic["stale"] if traffic else None,
                         "csrc_sha": sha,
                         "algorithmic_bytes_per_step_unfused": 256 * 54_600_000,
                         "kernel": f"trunk = stem_pool_kernel + bottleneck_chain_kernel ({n_chain} per pass: "
                                   f"layer1/layer2 conv2 -> conv3 [-> next conv1]) + conv_igemm[_cg2]_kernel; "
                                   f"{n_trunk} launches per step",
                         "trunk_ms_per_step": trunk_ms, "flop_per_frame": FLOP_PER_FRAME,
                         # the same FLOP count over the WHOLE step of the `value` region (uint8 frames in, K1's work done
                         # by the stem kernel's converter warps and not counted): the production path is faster than the
                         # trunk-only leg above, whose stem reads a 106 MB NHWC4p tensor instead of 38.5 MB of uint8
                         "whole_step": {"ms_per_step": ms_max / K,
                                        "achieved": BATCH * FLOP_PER_FRAME / (ms_max / K / 1e3) / 1e12,
                                        "frac": BATCH * FLOP_PER_FRAME / (ms_max / K / 1e3) / 1e12 / pk["bf16_burst"],
                                        "note": "per rank; at N > 1 the region also holds the NCCL gather"},
                         "frac_of_sustained_peak": achieved_tf / pk["bf16_sustained"],
                         "peak_sustained": pk["bf16_sustained"], "peak_source": pk["source"],
                         "sustained": {"achieved": sus_tf, "peak": pk["bf16_sustained"],
                                       "frac": sus_tf / pk["bf16_sustained"], "frac_of_burst_peak": sus_tf / pk["bf16_burst"],
                                       "trunk_ms_per_step": sus_trunk_ms, "replays": sus_trunk_n,
                                       "seconds": sus_trunk_ms * sus_trunk_n / 1e3, "clocks": sus_trunk_ck}},
            "sustained": {"value": world * BATCH / (sus_step_ms / 1e3), "unit": UNIT, "ms_per_step": sus_step_ms,
                          "steps": sus_step_n, "seconds": sus_step_ms * sus_step_n / 1e3, "clocks": sus_step_ck,
                          "note": "rank 0's own clock; whole steps (K1 + trunk), device-resident, no gather"},
            "roofline_k1": {"bound": "hbm", "achieved": k1_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": k1_gbs / pk["hbm_gbs"], "ms": k1_ms,
                            "bytes_per_frame": k1_bytes // BATCH,
                            "color_jitter_variant": {"ms": k1j_ms, "launches": 2,
                                                     "achieved": k1j_bytes / (k1j_ms / 1e3) / 1e9,
                                                     "frac": k1j_bytes / (k1j_ms / 1e3) / 1e9 / pk["hbm_gbs"],
                                                     "bytes_per_frame": k1j_bytes // BATCH}},
            "clocks": clocks.summary(),
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if gpu_ref:
            line["gpu_baseline"] = gpu_ref
        if cfg4:
            line["config4"] = cfg4
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
