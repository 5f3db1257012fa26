"""BASELINE config 3: batch-size sweep 1..1024 on one B200 — latency of the whole hot path (uint8 frames resident in
HBM -> features) and the implied tensor utilisation (2*MAC convention, measured peaks).

    python tools/bench_sweep.py > gpurun_out/sweep.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))
sys.path.insert(0, str(ROOT / "oracle"))

import torch  # noqa: E402

import phdfx  # noqa: E402
import resnet50_ref as R  # noqa: E402

FLOP_PER_FRAME = 2 * 4_087_136_256


def main():
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else \
        {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=1024)
    frames = torch.randint(0, 256, (2048, 224, 224, 3), dtype=torch.uint8, device="cuda")  # 308 MB > L2
    rows = []
    for b in [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]:
        out = torch.empty(b, 2048, device="cuda")
        n_slices = max(1, 2048 // b)
        for i in range(3):
            eng.extract_u8(frames[(i % n_slices) * b:(i % n_slices + 1) * b], None, out=out)
        torch.cuda.synchronize()
        reps = 20 if b <= 256 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            s = (i + 3) % n_slices
            eng.extract_u8(frames[s * b:(s + 1) * b], None, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fps = b / (ms / 1e3)
        tf = fps * FLOP_PER_FRAME / 1e12
        rows.append({"batch": b, "ms": round(ms, 4), "frames_per_s": round(fps, 1), "tflops": round(tf, 1),
                     "frac_of_burst_peak": round(tf / pk["bf16_tflops"], 4),
                     "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4)})
        print(f"batch {b:5d}  {ms:8.3f} ms  {fps:10.0f} frames/s  {tf:7.1f} TFLOP/s  "
              f"{100 * tf / pk['bf16_tflops_sustained']:5.1f}% of sustained peak", file=sys.stderr)
    print(json.dumps({"launches_per_call": eng.launches, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
