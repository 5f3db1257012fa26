"""Time each fused layer1 chain launch (bottleneck_chain_sm100.cuh) on its own buffers at a given batch, L2 flushed
between repetitions, next to the per-conv kernels it replaces.

    python tools/bench_chain.py [batch] [reps]
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402


def timed(fn, flush, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    for first in range(len(eng.plan.layers)):
        span = eng.chain_span(first)
        if span == 0:
            continue
        L3 = eng.plan.layers[first + 1]
        ds = L3.in2_buf >= 0
        hw, c2, n3 = L3.hin, L3.cin, L3.cout
        t1 = torch.relu(torch.randn(n, hw, hw, c2, device="cuda", generator=g)).to(torch.bfloat16)
        xr = torch.randn(n, hw, hw, 64 if ds else n3, device="cuda", generator=g).to(torch.bfloat16)
        ms = timed(lambda: eng.run_chain(first, t1, xr), flush, reps)
        # the per-conv kernels on the same data
        t2 = eng.run_layer(first, t1)
        parts = [timed(lambda: eng.run_layer(first, t1), flush, reps)]
        if ds:
            parts.append(timed(lambda: eng.run_layer(first + 1, t2, None, xr), flush, reps))
            out = eng.run_layer(first + 1, t2, None, xr)
        else:
            parts.append(timed(lambda: eng.run_layer(first + 1, t2, xr), flush, reps))
            out = eng.run_layer(first + 1, t2, xr)
        n1 = 0
        if span == 3:
            parts.append(timed(lambda: eng.run_layer(first + 2, out), flush, reps))
            n1 = eng.plan.layers[first + 2].cout
        px = n * hw * hw
        macs = px * (c2 * 9 * c2 + n3 * (c2 + (64 if ds else 0)) + n3 * n1)
        alg = px * 2 * (c2 + (64 if ds else n3) + n3 + n1)
        rows.append({"first": first, "name": "+".join(eng.plan.names[first:first + span]), "ms": round(ms, 4),
                     "per_conv_ms": [round(p, 4) for p in parts], "per_conv_sum_ms": round(sum(parts), 4),
                     "tflops": round(2 * macs / (ms / 1e3) / 1e12, 1), "alg_gbs": round(alg / (ms / 1e3) / 1e9, 1)})
        del t1, xr, t2, out
    print(json.dumps({"batch": n, "reps": reps, "chains": rows}, indent=1))
    for r in rows:
        print(f"{r['name']:60s} {r['ms']:8.4f} ms  (per-conv {r['per_conv_sum_ms']:.4f} = {r['per_conv_ms']})  "
              f"{r['tflops']:7.1f} TF/s  {r['alg_gbs']:7.1f} GB/s", file=sys.stderr)


if __name__ == "__main__":
    main()
