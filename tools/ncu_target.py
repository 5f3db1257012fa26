"""Small, deterministic launch sequences for ncu.

    python tools/ncu_target.py step [batch]          # 3 steps of the whole hot path from uint8 frames (40 launches each: K1 rides in the stem kernel)
    python tools/ncu_target.py trunk [batch]         # K1, then 3 passes of the trunk on its NHWC4p output (40 launches each) — what bench.py's roofline leg times
    python tools/ncu_target.py layers 0,3,5 [batch]  # each listed layer of the execution list twice via phdfx_run_layer
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402


def main():
    mode = sys.argv[1]
    if mode == "trunk":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
        frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
        x4 = eng.preprocess_u8(frames, None)
        for _ in range(3):
            eng.forward_nhwc4p(x4)
        torch.cuda.synchronize()
        print("trunk done, launches per pass:", eng.last_launch_count)
    elif mode == "step":
        n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
        eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
        if len(sys.argv) > 3:  # frame-wave schedule, e.g. 0:32,7:0 [reuse 0|1]
            eng.set_waves(tuple(tuple(int(v) for v in st.split(":")) for st in sys.argv[3].split(",")),
                          reuse=(sys.argv[4] != "0") if len(sys.argv) > 4 else True)
        frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda")
        for _ in range(3):
            eng.extract_u8(frames, None)
        torch.cuda.synchronize()
        print("step done, launches per step:", eng.launches)
    else:
        ids = [int(v) for v in sys.argv[2].split(",")]
        n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
        eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
        g = torch.Generator(device="cuda").manual_seed(0)
        for i in ids:
            L = eng.plan.layers[i]
            if L.kind in (1, 3):
                x = torch.zeros(n, 224, 232, 4, device="cuda", dtype=torch.bfloat16)
                x[:, :, 4:228, :3] = torch.randn(n, 224, 224, 3, device="cuda", generator=g).to(torch.bfloat16)
            else:
                x = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
            ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
            res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16) if L.res_buf >= 0 else None
            x2 = torch.randn(n, L.hin2, L.hin2, L.cin2, device="cuda", generator=g).to(torch.bfloat16) if L.in2_buf >= 0 else None
            for _ in range(2):
                eng.run_layer(i, x, res, x2)
            torch.cuda.synchronize()
            print("layer", i, eng.plan.names[i], "done")


if __name__ == "__main__":
    main()
