"""Debug: clock64 timeline of CTA 0 of one conv_igemm_kernel launch (PHDFX_CONV_TRACE; 1-CTA kernel only).

    python tools/trace_conv.py LAYER_ID [batch] > gpurun_out/conv_trace.txt
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
raw = str(ROOT / "gpurun_out" / "conv_trace_raw.txt")

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402

EV = {0: "tma first load", 1: "tma last load", 2: "mma tile start", 3: "mma acc free", 4: "mma issued", 8: "dma res g0",
      9: "dma res g1", 10: "dma res g2", 11: "dma res g3", 12: "dma st g0", 13: "dma st g1", 14: "dma st g2",
      15: "dma st g3", 16: "epi tile start", 17: "epi acc full", 18: "epi buf g0", 19: "epi buf g1", 20: "epi buf g2",
      21: "epi buf g3", 22: "epi done g0", 23: "epi done g1", 24: "epi done g2", 25: "epi done g3"}
i = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
L = eng.plan.layers[i]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16) if L.res_buf >= 0 else None
eng.run_layer(i, x, res)
torch.cuda.synchronize()
os.environ["PHDFX_CONV_TRACE"] = raw
eng.run_layer(i, x, res)
torch.cuda.synchronize()
rows = [[int(v) for v in ln.split()[1:]] for ln in open(raw).read().strip().splitlines()]
t0 = min(v for r in rows for v in r if v > 0)
print("layer", i, eng.plan.names[i])
ev = sorted((v - t0, k, e) for k, r in enumerate(rows) for e, v in enumerate(r) if v > 0 and 3 <= k <= 5)
for t, k, e in ev:
    print(f"{t:8d}  tile {k}  {EV.get(e, e)}")
st = [r[17] for r in rows if r[17] > 0]
print("epilogue tile period:", [st[j + 1] - st[j] for j in range(2, min(len(st) - 1, 10))])
