"""K1 inside the stem kernel vs K1 as a launch of its own, whole step under graph replay, for a frame geometry.

    python tools/ab_fuse_k1.py N H W [BOX_SIDE]
"""
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
import torch
import phdfx
from phdfx import synthetic as R

n, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
side = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bb = R.seeded_backbone()
frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
boxes = None
if side:
    boxes = torch.tensor([((H - side) // 2, (W - side) // 3, side, side)] * n, dtype=torch.int32, device="cuda")
res = {}
outs = []
graphs = []
for name, env in (("fused", {}), ("unfused", {"PHDFX_NO_FUSE_K1": "1"})):
    os.environ.update(env)
    try:
        eng = phdfx.B200Backbone(bb, device=0, max_frames=n)
    finally:
        for k in env:
            del os.environ[k]
    out = torch.empty(n, 2048, device="cuda")
    for _ in range(3):
        eng.extract_u8(frames, boxes, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        eng.extract_u8(frames, boxes, out=out)
    torch.cuda.current_stream().wait_stream(st)
    with torch.cuda.graph(g):
        eng.extract_u8(frames, boxes, out=out)
    graphs.append((name, g, eng))
    outs.append(out)
    res[name] = []
for r in range(8):
    for name, g, _ in graphs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        if r:
            res[name].append(e0.elapsed_time(e1) / 30 * 1e3)
print(f"n {n}, frames {H}x{W}, box side {side or 'none'}: " +
      ", ".join(f"{k} median {statistics.median(v):.1f} us/step (min {min(v):.1f})" for k, v in res.items()) +
      f", bit-identical {torch.equal(outs[0], outs[1])}")
