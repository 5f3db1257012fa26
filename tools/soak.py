"""Soak: many passes of the default path (K1 in the stem, multi-phase launches, chains) at several batch sizes, eager and
under graph replay, every result compared bit for bit with a handle that uses none of the cross-launch / cross-phase
synchronisation (PHDFX_NO_MULTI=1, PHDFX_NO_FUSE_K1=1).  A missed dependency would show up as a rare mismatch.

    python tools/soak.py [iterations]
"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
import torch
import phdfx
from phdfx import synthetic as R

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bb = R.seeded_backbone()
bad = 0
for n in (256, 200, 131, 97):
    eng = phdfx.B200Backbone(bb, device=0, max_frames=n)
    os.environ["PHDFX_NO_MULTI"] = "1"
    os.environ["PHDFX_NO_FUSE_K1"] = "1"
    try:
        ref_eng = phdfx.B200Backbone(bb, device=0, max_frames=n)
    finally:
        del os.environ["PHDFX_NO_MULTI"], os.environ["PHDFX_NO_FUSE_K1"]
    inputs = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda") for _ in range(3)]
    wants = [ref_eng.extract_u8(x, None).clone() for x in inputs]
    out = torch.empty(n, 2048, device="cuda")
    graphs = [eng.capture_extract(x, None) for x in inputs]
    for it in range(iters):
        j = it % 3
        out.fill_(float("nan"))
        eng.extract_u8(inputs[j], None, out=out)
        if not torch.equal(out, wants[j]):
            bad += 1
            print(f"n {n} iteration {it}: eager mismatch, max diff {(out - wants[j]).abs().max().item()}")
        if not torch.equal(graphs[j].replay(), wants[j]):
            bad += 1
            print(f"n {n} iteration {it}: graph mismatch")
    print(f"n {n}: {iters} eager + {iters} replayed passes, launches {eng.launches} vs {ref_eng.launches}, mismatches so far {bad}")
    del graphs
    eng.close()
    ref_eng.close()
print("SOAK", "FAILED" if bad else "OK")
sys.exit(1 if bad else 0)
