"""Time one layer of the execution list at a batch (CUDA events, L2 flushed): python tools/time_layer.py LAYER [batch] [reps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200")); import torch, phdfx
from phdfx import synthetic as R
i = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 256; reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
L = eng.plan.layers[i]
g = torch.Generator(device="cuda").manual_seed(0)
if L.kind in (1, 3):
    x = torch.zeros(n, 224, 232, 4, device="cuda", dtype=torch.bfloat16)
    x[:, :, 4:228, :3] = torch.randn(n, 224, 224, 3, device="cuda", generator=g).to(torch.bfloat16)
else:
    x = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16) if L.res_buf >= 0 else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
x2 = torch.randn(n, L.hin2, L.hin2, L.cin2, device="cuda", generator=g).to(torch.bfloat16) if L.in2_buf >= 0 else None
eng.run_layer(i, x, res, x2); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run_layer(i, x, res, x2); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"layer {i} {eng.plan.names[i]} batch {n}: median {sorted(ts)[len(ts)//2]*1e3:.1f} us  min {min(ts)*1e3:.1f} us")
