"""Debug: clock64 timeline of CTA 0 of each fused layer1 chain (PHDFX_CHAIN_TRACE), decoded relative to tile 0.

    python tools/trace_chain.py [batch] > gpurun_out/chain_trace.txt
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
raw = str(ROOT / "gpurun_out" / "chain_trace_raw.txt")
if os.path.exists(raw):
    os.remove(raw)
os.environ["PHDFX_CHAIN_TRACE"] = raw

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402

EV = {0: "mma conv2 start", 1: "mma conv2 issued", 2: "mma conv3 issue", 3: "mma conv1n start", 4: "mma conv1n issued",
      5: "epi A start", 6: "epi A end", 7: "epi B start", 8: "epi B g0", 9: "epi B g1", 10: "epi B g2", 11: "epi B g3",
      12: "epi C start", 13: "epi C end", 14: "epi B.hi start", 15: "halo load issue", 16: "dma S0", 17: "dma S1",
      18: "dma S2", 19: "dma S3", 20: "dma S4", 21: "dma S5", 22: "res load B0 issue"}

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
g = torch.Generator(device="cuda").manual_seed(0)
for first in range(len(eng.plan.layers)):
    if eng.chain_span(first) == 0:
        continue
    L3 = eng.plan.layers[first + 1]
    ds = L3.in2_buf >= 0
    t1 = torch.relu(torch.randn(n, L3.hin, L3.hin, L3.cin, device="cuda", generator=g)).to(torch.bfloat16)
    xr = torch.randn(n, L3.hin, L3.hin, 64 if ds else L3.cout, device="cuda", generator=g).to(torch.bfloat16)
    eng.run_chain(first, t1, xr)
    torch.cuda.synchronize()
blocks = open(raw).read().strip().split("chain ")[1:]
for blk in blocks:
    lines = blk.strip().splitlines()
    print("== chain", lines[0])
    rows = [[int(v) for v in ln.split()[1:]] for ln in lines[1:]]
    t0 = min(v for r in rows for v in r if v > 0)
    events = sorted((v - t0, k, e) for k, r in enumerate(rows) for e, v in enumerate(r) if v > 0 and k < 6)
    for t, k, e in events:
        print(f"{t:8d}  tile {k}  {EV.get(e, e)}")
    # steady-state period
    c3 = [r[2] for r in rows if r[2] > 0]
    if len(c3) > 8:
        print("conv3 issue period (cycles) tiles 4..:", [c3[i + 1] - c3[i] for i in range(4, min(len(c3) - 1, 16))])
