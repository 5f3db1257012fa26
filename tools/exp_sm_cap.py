import os, sys, json
sys.path.insert(0, "implementation-phd-lab-vision_b200"); sys.path.insert(0, "oracle")
import torch, phdfx, resnet50_ref as R
cap = os.environ.get("PHDFX_SM_CAP", "148")
n = 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
g = torch.Generator(device="cuda").manual_seed(0)
for i in [27, 25, 26, 13, 45]:
    L = eng.plan.layers[i]
    x = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
    ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
    res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16) if L.res_buf >= 0 else None
    eng.run_layer(i, x, res); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.run_layer(i, x, res); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"cap {cap:>4s} layer {i} {eng.plan.names[i]:16s} {sorted(ts)[2]:.4f} ms")
