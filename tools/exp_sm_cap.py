"""Experiment: how a layer's time changes when every persistent grid runs on fewer SMs (PHDFX_SM_CAP=n, read at
phdfx_create).  A layer bound by something every SM owns privately (tensor pipe, shared memory, its L2 port) slows down
in proportion to the SMs taken away; one bound by a shared resource (HBM, aggregate L2 bandwidth) does not.

    for c in 148 110 74 36; do PHDFX_SM_CAP=$c python tools/exp_sm_cap.py; done

Measured on the v8 build (B200, batch 256, L2 flushed): layer3.1.conv3 0.051 / 0.059 / 0.074 / 0.129 ms at
148 / 110 / 74 / 36 SMs, layer3.1.conv2 (tensor-bound 3x3) 0.051 / 0.062 / 0.082 / 0.133 — half the SMs cost 1.4-1.6x,
not 2x: at 148 SMs both are partly limited by what the SMs share (aggregate L2 -> SM operand bandwidth).
"""
import os, sys, json
sys.path.insert(0, "implementation-phd-lab-vision_b200")
import torch, phdfx
from phdfx import synthetic as R
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
