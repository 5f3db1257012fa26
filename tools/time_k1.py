"""K1 alone at batch 256: plain and colour-jitter variant, identity-size and cropped boxes (CUDA events, 20 reps)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200")); import torch, phdfx
from phdfx import synthetic as R
n = 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
jit = torch.stack([phdfx.jitter_params([2, 0, 3, 1], 1.2, 0.9, 1.1, 0.02)] * n).cuda()
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, H, W, box in (("identity 224x224", 224, 224, None), ("crop 241x241 of 260x300", 260, 300, (5, 7, 241, 241))):
    frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    boxes = None if box is None else torch.tensor([box] * n, dtype=torch.int32, device="cuda")
    out = eng.preprocess_u8(frames, boxes)
    t_plain = timeit(lambda: eng.preprocess_u8(frames, boxes, out=out))
    t_jit = timeit(lambda: eng.preprocess_u8(frames, boxes, out=out, jitter=jit))
    print(f"{name}: plain {t_plain*1e3:.1f} us, colour-jitter variant {t_jit*1e3:.1f} us (2 launches)")
