"""Generates tests/golden/*.npz.  Runs ONLY in the build container: it imports the reference
(/root/reference/src/dataset.py through a one-line VideoReader shim — torchvision.io.VideoReader no longer exists)
and torchvision's CPU fp32 ResNet-50, i.e. the reference's own arithmetic, and stores their outputs as golden vectors.

    python oracle/make_golden.py

Fixtures
  preprocess_golden.npz   reference `_crop_and_resize_video_uint8` (src/dataset.py:141-152) + Normalize (:242-245)
                          outputs for seeded uint8 frames, under both ATen kernel choices (1 thread = the DataLoader
                          worker path; N threads = the generic path)
  trunk_golden.npz        reference trunk (src/preprocess_resnet_features.py:207-209,296; torchvision CPU fp32)
                          features for the config-1 stand-in: 16 frames = 2 clips x 8, 1002x1000 uint8, box
                          (100,200,517,517), plus per-stage activation statistics
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import torchvision.io as tio  # noqa: E402

tio.VideoReader = None  # shim: the symbol was removed upstream; the functions we call never touch it
sys.path.insert(0, "/root/reference/src")
import dataset as refds  # noqa: E402  (the reference, unmodified)

import resnet50_ref as R  # noqa: E402

OUT = ROOT / "tests" / "golden"
OUT.mkdir(parents=True, exist_ok=True)

PRE_CASES = [  # (n, H, W, box(top,left,h,w), seed)
    (2, 300, 280, (20, 30, 217, 217), 11),   # downscale, odd box
    (1, 224, 224, (0, 0, 224, 224), 12),     # identity early-out (functional.py:470-471)
    (1, 160, 200, (5, 9, 64, 64), 13),       # 3.5x upscale: many exact .5 ties
    (1, 400, 420, (3, 7, 113, 113), 14),     # upscale, prime side
    (1, 600, 640, (37, 41, 517, 517), 15),   # H36M-like box side
    (1, 90, 90, (0, 0, 1, 1), 16),           # degenerate 1x1 box (side_i clamps to >= 1, dataset.py:103)
]


def ref_preprocess(frames_u8: np.ndarray, box, threads: int):
    torch.set_num_threads(threads)
    fr = torch.from_numpy(frames_u8)
    video = refds._crop_and_resize_video_uint8(fr, torch.tensor(box, dtype=torch.int64), out_size=224)
    tf = refds.T.Compose([refds.T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))])
    return video.numpy(), tf(video).numpy()


def main():
    nthreads = torch.get_num_threads()
    # ---------------- preprocess
    pre = {}
    for i, (n, H, W, box, seed) in enumerate(PRE_CASES):
        frames = R.seeded_frames(n, H, W, seed)
        v1, x1 = ref_preprocess(frames, box, 1)
        vN, xN = ref_preprocess(frames, box, max(2, nthreads))
        pre[f"case{i}_meta"] = np.array([n, H, W, *box, seed], dtype=np.int64)
        pre[f"case{i}_u8_worker"] = np.round(v1 * 255).astype(np.uint8)
        pre[f"case{i}_u8_generic"] = np.round(vN * 255).astype(np.uint8)
        # normalised fp32 is a pure function of the uint8 image; keep one full copy for the /255, (x-mean)/std check
        if i == 0:
            pre["case0_norm_worker"] = x1.astype(np.float32)
    np.savez_compressed(OUT / "preprocess_golden.npz", **pre)
    torch.set_num_threads(nthreads)

    # ---------------- trunk (config 1 stand-in)
    backbone = R.seeded_backbone()
    frames = R.seeded_frames(16, 1002, 1000, 1)
    box = (100, 200, 517, 517)
    _, x = ref_preprocess(frames, box, 1)  # (16,3,224,224) fp32 — the tensor the reference feeds to backbone()
    torch.set_num_threads(nthreads)
    xt = torch.from_numpy(x)
    taps = {}
    with torch.no_grad():
        h = xt
        for idx, m in enumerate(backbone):
            h = m(h)
            if idx in (3, 4, 5, 6, 7):
                taps[idx] = h
        feats = h.flatten(1)  # :296
    trunk = {
        "feats": feats.numpy().astype(np.float32),
        "meta": np.array([16, 1002, 1000, *box, 1, R.WEIGHT_SEED, R.BN_SEED], dtype=np.int64),
        "x_checksum": np.array([float(xt.double().sum()), float(xt.double().abs().sum())]),
    }
    for idx, t in taps.items():
        trunk[f"tap{idx}_absmean"] = np.array(float(t.double().abs().mean()))
        trunk[f"tap{idx}_max"] = np.array(float(t.max()))
        trunk[f"tap{idx}_frame0_ch0"] = t[0, 0].numpy().astype(np.float32)
    # identity-size input (224x224, no resize) for 4 frames
    frames2 = R.seeded_frames(4, 224, 224, 2)
    _, x2 = ref_preprocess(frames2, (0, 0, 224, 224), 1)
    torch.set_num_threads(nthreads)
    with torch.no_grad():
        trunk["feats_identity"] = backbone(torch.from_numpy(x2)).flatten(1).numpy().astype(np.float32)
    np.savez_compressed(OUT / "trunk_golden.npz", **trunk)
    print("wrote", [p.name for p in OUT.iterdir()])
    print("feats absmean", float(np.abs(trunk["feats"]).mean()), "max", float(trunk["feats"].max()))


if __name__ == "__main__":
    main()
