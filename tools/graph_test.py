"""Eager launches vs CUDA-graph replay of one whole step (K1 + trunk) at batch 256."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200")); import torch, phdfx
from phdfx import synthetic as R
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(n, 2048, device="cuda")
for _ in range(3):
    eng.extract_u8(frames, None, out=out)
torch.cuda.synchronize()
def timeit(fn, reps=100):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
eager = timeit(lambda: eng.extract_u8(frames, None, out=out))
ref = out.clone()
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    eng.extract_u8(frames, None, out=out)
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    eng.extract_u8(frames, None, out=out)
out.zero_()
g.replay(); torch.cuda.synchronize()
print("graph result equal:", torch.equal(out, ref))
graph = timeit(g.replay)
print(f"batch {n}: eager {eager*1e3:.1f} us/step, graph replay {graph*1e3:.1f} us/step")
