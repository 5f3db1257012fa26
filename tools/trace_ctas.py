"""CTA timeline of one pass at batch N (PHDFX_CTA_TRACE): for every conv launch of layer3 / layer4, when its CTAs became
resident, when their first tile's inputs were available, when they exited — relative to the previous launch's last exit.

    PHDFX_EXPERIMENTAL=1 python -c "import __graft_entry__ as g; g.build(force=True)"     (the timeline code is not in a normal build)
    python tools/trace_ctas.py [batch] > gpurun_out/cta_trace.txt        (PHDFX_FLAGS=1: with frame progress links)
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
raw = str(ROOT / "gpurun_out" / "cta_trace_raw.txt")
os.makedirs(os.path.dirname(raw), exist_ok=True)
os.environ["PHDFX_CTA_TRACE"] = raw
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
import numpy as np
import torch
import phdfx
from phdfx import synthetic as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(n, 2048, device="cuda")
for _ in range(4):
    eng.extract_u8(frames, None, out=out)
torch.cuda.synchronize()
rows = np.loadtxt(raw, dtype=np.int64)
names = eng.launch_names() if hasattr(eng, "launch_names") else None
t_base = rows[:, 2][rows[:, 2] > 0].min()
prev_end = None
print(f"batch {n}, links {eng.linked_launches(n)}; times in us; 'idle' = sum over CTAs of (first inputs - resident), "
      f"'depwait' = sum over CTAs of time the TMA producer warp waited on counters")
print(f"{'layer':>5} {'link':>4} {'ctas':>4} {'first_res':>9} {'last_res':>9} {'first_in':>9} {'last_in':>9} {'first_exit':>10} "
      f"{'last_exit':>9} {'dur':>7} {'gap_prev':>8} {'idle_us':>8} {'depwait':>8} {'first_entry':>11} {'prologue':>8}")
for l in sorted(set(rows[:, 0])):
    r = rows[rows[:, 0] == l]
    res, tin, ex, dw = (r[:, 2] - t_base) / 1e3, (r[:, 3] - t_base) / 1e3, (r[:, 4] - t_base) / 1e3, r[:, 5] / 1e3
    ent = (r[:, 7] - t_base) / 1e3
    gap = "" if prev_end is None else f"{res.min() - prev_end:8.1f}"
    print(f"{l:5d} {r[0, 6]:4d} {len(r):4d} {res.min():9.1f} {res.max():9.1f} {tin.min():9.1f} {tin.max():9.1f} {ex.min():10.1f} "
          f"{ex.max():9.1f} {ex.max() - res.min():7.1f} {gap:>8} {(tin - res).sum():8.0f} {dw.sum():8.0f} {ent.min():11.1f} "
          f"{np.median(res - ent):8.2f}")
    prev_end = ex.max()
