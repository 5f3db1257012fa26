"""BASELINE config 3: batch-size sweep 1..1024 on one B200 — latency of the whole hot path (uint8 frames resident in
HBM -> features), under CUDA-graph replay (the latency a caller of ExtractGraph sees) and with plain launches, and per
launch of the trunk (phdfx_forward_timed: CUDA events between launches, in situ) time + computed tensor utilisation
2*MAC*N / (t * peak) against the measured dense-bf16 peaks.

    python tools/bench_sweep.py > gpurun_out/sweep.json          (table on stderr)
"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402

import phdfx  # noqa: E402
from phdfx import synthetic as R  # noqa: E402

FLOP_PER_FRAME = 2 * 4_087_136_256
BATCHES = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024]


def launch_macs(eng):
    """MACs per frame of every launch of the un-waved list (a fused chain = the sum of its convs)."""
    per_layer = []
    for L in eng.plan.layers:
        if L.kind == 2:
            per_layer.append(0)
            continue
        ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
        m = ho * ho * L.cout * L.cin * L.r * L.s
        if L.in2_buf >= 0:
            m += ho * ho * L.cout * L.cin2
        per_layer.append(m)
    out, i = [], 0
    while i < len(per_layer):
        span = max(1, eng.chain_span(i))
        out.append(sum(per_layer[i:i + span]))
        i += span
    return out


def timeit(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else \
        {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=1024)
    frames = torch.randint(0, 256, (2048, 224, 224, 3), dtype=torch.uint8, device="cuda")  # 308 MB > L2
    macs = launch_macs(eng)
    rows = []
    for b in BATCHES:
        n_slices = max(1, 2048 // b)
        n_g = min(n_slices, 8)
        outs = [torch.empty(b, 2048, device="cuda") for _ in range(n_g)]
        graphs = [eng.capture_extract(frames[s * b:(s + 1) * b], None, out=outs[s]) for s in range(n_g)]
        for g in graphs:
            g.replay()
        torch.cuda.synchronize()
        reps = 40 if b <= 64 else 16 if b <= 256 else 8
        ms_graph = min(timeit(lambda i: graphs[i % n_g].replay(), reps) for _ in range(3))
        ms_eager = timeit(lambda i: eng.extract_u8(frames[(i % n_slices) * b:(i % n_slices + 1) * b], None,
                                                   out=outs[0]), reps)
        # per launch, in situ (median of 5)
        x4 = eng.preprocess_u8(frames[:b], None)
        runs = [eng.forward_timed(x4)[1] for _ in range(5)]
        names = [nm for nm, _ in runs[0]]
        med = [sorted(r[i][1] for r in runs)[2] for i in range(len(names))]
        per_launch = [{"name": nm, "ms": round(ms, 5),
                       "tflops": round(2 * mc * b / (ms / 1e3) / 1e12, 1) if ms > 0 else 0.0,
                       "frac_of_burst_peak": round(2 * mc * b / (ms / 1e3) / 1e12 / pk["bf16_tflops"], 4) if ms > 0 else 0.0}
                      for nm, ms, mc in zip(names, med, macs)]
        fps = b / (ms_graph / 1e3)
        tf = fps * FLOP_PER_FRAME / 1e12
        rows.append({"batch": b, "ms_graph_replay": round(ms_graph, 4), "ms_plain_launches": round(ms_eager, 4),
                     "frames_per_s": round(fps, 1), "tflops": round(tf, 1),
                     "frac_of_burst_peak": round(tf / pk["bf16_tflops"], 4),
                     "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4),
                     "launches": graphs[0].launches, "trunk_insitu_sum_ms": round(sum(med), 4),
                     "per_launch": per_launch})
        print(f"batch {b:5d}  graph {ms_graph:8.4f} ms  plain {ms_eager:8.4f} ms  {fps:10.0f} frames/s  {tf:7.1f} TFLOP/s  "
              f"{100 * tf / pk['bf16_tflops']:5.1f}% of burst peak  ({graphs[0].launches} launches)", file=sys.stderr)
        del graphs
    # per-launch utilisation table: one row per launch, one column per batch
    print("\n% of burst bf16 peak per launch (in situ), by batch " + " ".join(f"{b:>6d}" for b in BATCHES),
          file=sys.stderr)
    for li, nm in enumerate([p["name"] for p in rows[0]["per_launch"]]):
        cells = " ".join(f"{100 * r['per_launch'][li]['frac_of_burst_peak']:6.1f}" for r in rows)
        print(f"{nm[:52]:52s} {cells}", file=sys.stderr)
    print(json.dumps({"peaks": {"burst": pk["bf16_tflops"], "sustained": pk["bf16_tflops_sustained"]}, "rows": rows},
                     indent=1))


if __name__ == "__main__":
    main()
