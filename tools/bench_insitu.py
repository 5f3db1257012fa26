"""In-situ per-launch timing of one trunk pass (phdfx_forward_timed): every launch runs with the L2 contents its
predecessor left, unlike tools/bench_layers.py (isolated layers, L2 flushed).  Median over repetitions.

    python tools/bench_insitu.py [batch] [reps] [waves, e.g. 0:32,7:0] [reuse 0|1] > gpurun_out/insitu.json
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
    if len(sys.argv) > 3:
        eng.set_waves(tuple(tuple(int(v) for v in st.split(":")) for st in sys.argv[3].split(",")),
                      reuse=(sys.argv[4] != "0") if len(sys.argv) > 4 else True)
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [eng.preprocess_u8(torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device="cuda", generator=g))
          for _ in range(3)]
    for x in xs:
        eng.forward_nhwc4p(x)
    runs = []
    for r in range(reps):
        _, t = eng.forward_timed(xs[r % 3])
        runs.append(t)
    names = [nm for nm, _ in runs[0]]
    med = [sorted(run[i][1] for run in runs)[reps // 2] for i in range(len(names))]
    # MACs per launch for the utilisation column
    macs = {}
    for nm, L in zip(eng.plan.names, eng.plan.layers):
        if L.kind == 2:
            macs[nm] = 0
            continue
        ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
        m = n * ho * ho * L.cout * L.cin * L.r * L.s
        if L.in2_buf >= 0:
            m += n * ho * ho * L.cout * L.cin2
        macs[nm] = m
    rows = []
    for nm, ms in zip(names, med):
        m = sum(macs[part] for part in _split(nm, macs))
        rows.append({"name": nm, "ms": round(ms, 4), "tflops": round(2 * m / (ms / 1e3) / 1e12, 1) if ms > 0 else 0})
    tot = sum(med)
    print(json.dumps({"batch": n, "reps": reps, "sum_ms": tot, "launches": rows}, indent=1))
    for r in rows:
        print(f"{r['name']:64s} {r['ms']:8.4f} ms {100 * r['ms'] / tot:5.1f}%  {r['tflops']:7.1f} TF/s", file=sys.stderr)
    print(f"sum {tot:.3f} ms over {len(rows)} launches (events serialise the launches)", file=sys.stderr)


def _split(name, macs):
    """'a.conv2+a.conv3+downsample+b.conv1' -> plan names (a plan name may itself contain '+')."""
    parts, cur = [], ""
    for tok in name.split("+"):
        cur = tok if not cur else cur + "+" + tok
        if cur in macs and not (cur + "+downsample" in macs and name.find(cur + "+downsample") >= 0) and \
                not (cur + "+maxpool" in macs):
            parts.append(cur)
            cur = ""
    assert not cur, (name, cur)
    return parts


if __name__ == "__main__":
    main()
