"""Frame-wave schedules at batch 256: bit-identity against the un-waved pass and time per step under graph replay.

    python tools/exp_waves.py [batch] [reps] > gpurun_out/exp_waves.json

Every schedule is (after_block, frames_per_wave) pairs as B200Backbone.set_waves takes them, with and without
PHDFX_SCHED_REUSE.  The timed region replays `reps` graphs of one whole step (K1 + trunk, uint8 frames in HBM) cycling
over 4 input batches (4 x 38.5 MB uint8), so nothing survives in L2 from one step to the next except by design.
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402

SCHEDULES = [
    ((0, 0),),
    ((0, 16), (7, 0)),
    ((0, 21), (7, 0)),
    ((0, 24), (7, 0)),
    ((0, 32), (7, 0)),
    ((0, 37), (7, 0)),
    ((0, 43), (7, 0)),
    ((0, 64), (7, 0)),
    ((0, 128), (7, 0)),
    ((0, 21), (3, 42), (7, 0)),
    ((0, 32), (3, 64), (7, 0)),
    ((0, 37), (3, 74), (7, 0)),
    ((0, 32), (3, 0)),
    ((0, 37), (3, 0)),
    ((0, 32), (7, 128), (13, 0)),
    ((0, 37), (7, 128), (13, 0)),
    ((0, 32), (7, 128)),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    only = sys.argv[3] if len(sys.argv) > 3 else None
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
    g = torch.Generator(device="cuda").manual_seed(5)
    batches = [torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
    outs = [torch.empty(n, 2048, device="cuda") for _ in range(4)]
    eng.set_waves(((0, 0),))
    ref = [eng.extract_u8(b).clone() for b in batches]
    rows = []
    scheds = SCHEDULES if only is None else [tuple(tuple(int(v) for v in st.split(":")) for st in only.split(","))]
    for sched in scheds:
        for reuse in (True, False):
            if sched == ((0, 0),) and not reuse:
                continue
            try:
                eng.set_waves(sched, reuse=reuse)
            except RuntimeError as e:  # noqa: PERF203
                rows.append({"waves": sched, "reuse": reuse, "error": str(e)})
                print(rows[-1], file=sys.stderr)
                continue
            graphs = [eng.capture_extract(batches[i], None, out=outs[i]) for i in range(4)]
            for i in range(4):
                outs[i].zero_()
                graphs[i].replay()
            torch.cuda.synchronize()
            equal = all(torch.equal(outs[i], ref[i]) for i in range(4))
            best = None
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(reps):
                    graphs[i % 4].replay()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                best = ms if best is None else min(best, ms)
            rows.append({"waves": sched, "reuse": reuse, "bit_identical": equal, "ms_per_step": round(best, 4),
                         "frames_per_s": round(n / best * 1e3), "launches": graphs[0].launches})
            print(rows[-1], file=sys.stderr)
            del graphs
    print(json.dumps({"batch": n, "reps": reps, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
