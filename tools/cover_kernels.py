"""A short eager pass over every kernel family, chains on and off, with a bit-equality check between the two.

    python tools/cover_kernels.py [n_frames]

Runs the whole hot path (no CUDA graph) on a ragged batch with real crop boxes, the colour-jitter variant of K1, the
per-launch timed forward, and the same batch with the bottleneck chains switched off (so the per-conv HALO / TILED
kernels the chains replace are covered as well).  Small enough to sit under a memory checker where one is available
(compute-sanitizer is closed on this pool).
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402


def run(n, chains):
    if chains:
        os.environ.pop("PHDFX_NO_CHAIN", None)
    else:
        os.environ["PHDFX_NO_CHAIN"] = "1"
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
    g = torch.Generator().manual_seed(3)
    frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, generator=g).cuda()
    boxes = torch.tensor([[10 + 3 * i, 7 + 2 * i, 200 - i, 190 - 2 * i] for i in range(n)], dtype=torch.int32).cuda()
    outs = [eng.extract_u8(frames, None).clone(), eng.extract_u8(frames, boxes).clone()]
    jit = torch.stack([phdfx.jitter_params([2, 0, 3, 1], 1.2 - 0.02 * i, 0.9, 1.1, 0.02) for i in range(n)]).cuda()
    outs.append(eng.extract_u8(frames, boxes, jitter=jit).clone())
    x = eng.preprocess_u8(frames, boxes)
    feats, times = eng.forward_timed(x)
    outs.append(feats.clone())
    torch.cuda.synchronize()
    print(f"chains={chains}: {eng.last_launch_count} launches in the last call, {len(times)} timed, finite={all(torch.isfinite(o).all().item() for o in outs)}")
    return outs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    a = run(n, True)
    b = run(n, False)
    same = all(torch.equal(x, y) for x, y in zip(a, b))
    print("chains vs per-conv launches bit-identical:", same)
    if not same:
        sys.exit(3)


if __name__ == "__main__":
    main()
