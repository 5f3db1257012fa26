"""In-process A/B of per-handle switches: one engine per environment setting (the switches are read at phdfx_create),
one captured graph of a whole step (K1 + trunk) each, replays interleaved round by round on the same GPU at the same
clocks.  Prints the median / min time per step of every setting and checks the features are bit-identical.

    python tools/ab_handles.py BATCH ROUNDS VAR=val[,VAR2=val2] [VAR=val ...]      ("-" = default environment)
    python tools/ab_handles.py 256 7 - PHDFX_FLAGS=1 PHDFX_FLAGS=1,PHDFX_FLAG_MAX_MB=1000
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

n, rounds = int(sys.argv[1]), int(sys.argv[2])
settings = sys.argv[3:] or ["-", "PHDFX_FLAGS=1"]
bb = R.seeded_backbone()
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
engines, graphs, outs = [], [], []
for s in settings:
    kv = {} if s == "-" else dict(x.split("=") for x in s.split(","))
    os.environ.update(kv)
    try:
        eng = phdfx.B200Backbone(bb, device=0, max_frames=n)
    finally:
        for k in kv:
            del os.environ[k]
    out = torch.empty(n, 2048, device="cuda")
    for _ in range(3):
        eng.extract_u8(frames, None, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        eng.extract_u8(frames, None, out=out)
    torch.cuda.current_stream().wait_stream(st)
    with torch.cuda.graph(g):
        eng.extract_u8(frames, None, out=out)
    engines.append(eng), graphs.append(g), outs.append(out)
reps = 40
times = [[] for _ in settings]
for r in range(rounds + 1):
    for i, g in enumerate(graphs):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        if r > 0:  # round 0 warms up
            times[i].append(e0.elapsed_time(e1) / reps * 1e3)
for i, s in enumerate(settings):
    same = torch.equal(outs[i], outs[0])
    print(f"{s:40s} links {engines[i].linked_launches(n):2d}  median {statistics.median(times[i]):8.1f} us/step  "
          f"min {min(times[i]):8.1f}  bit-identical to the first: {same}")
