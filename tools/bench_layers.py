"""Per-layer timing of the trunk at a given batch (BASELINE config 3 building block): each entry of the execution
list is run through phdfx_run_layer on its own buffers, CUDA-event timed, with an L2 flush between repetitions.

    python tools/bench_layers.py [batch] [reps] > gpurun_out/layers.json

Reports, per layer: time, TFLOP/s (2*MAC) and fraction of the measured bf16 peak, algorithmic bytes (activations in +
out + residual + weights) and GB/s vs the measured HBM peak, and which roofline bounds it.
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
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else \
        {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    eng = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []
    tot_ms = 0.0
    for i, (name, L) in enumerate(zip(eng.plan.names, eng.plan.layers)):
        if L.kind in (1, 3):
            x = torch.zeros(n, 224, 232, 4, device="cuda", dtype=torch.bfloat16)
            x[:, :, 4:228, :3] = torch.randn(n, 224, 224, 3, device="cuda", generator=g).to(torch.bfloat16)
            in_bytes = n * 224 * 232 * 4 * 2
        else:
            x = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
            in_bytes = x.numel() * 2
        ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
        res = None
        if L.res_buf >= 0:
            res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16)
        out_bytes = n * ho * ho * L.cout * 2 if not L.gap else n * L.cout * 4
        if L.kind == 3:
            out_bytes = n * 56 * 56 * 64 * 2
        w_bytes = 0 if L.kind == 2 else (L.r * L.s * L.cin * L.cout * 2)
        macs = 0 if L.kind == 2 else n * ho * ho * L.cout * L.cin * L.r * L.s
        alg_bytes = in_bytes + out_bytes + (res.numel() * 2 if res is not None else 0) + w_bytes
        x2 = None
        if L.in2_buf >= 0:
            x2 = torch.randn(n, L.hin2, L.hin2, L.cin2, device="cuda", generator=g).to(torch.bfloat16)
            in_bytes += x2.numel() * 2
            w_bytes += L.cin2 * L.cout * 2
            macs += n * ho * ho * L.cout * L.cin2
            alg_bytes = in_bytes + out_bytes + w_bytes
        eng.run_layer(i, x, res, x2)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.run_layer(i, x, res, x2)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tot_ms += ms
        tf = 2 * macs / (ms / 1e3) / 1e12
        gbs = alg_bytes / (ms / 1e3) / 1e9
        t_tensor = 2 * macs / (pk["bf16_tflops"] * 1e12) * 1e3
        t_hbm = alg_bytes / (pk["hbm_gbs"] * 1e9) * 1e3
        rows.append({"i": i, "name": name, "cin": L.cin, "cout": L.cout, "r": L.r, "stride": L.stride, "hin": L.hin,
                     "res": L.res_buf >= 0, "ms": round(ms, 4), "tflops": round(tf, 1),
                     "frac_tensor_burst": round(tf / pk["bf16_tflops"], 3), "gbs": round(gbs, 1),
                     "frac_hbm": round(gbs / pk["hbm_gbs"], 3), "ideal_ms": round(max(t_tensor, t_hbm), 4),
                     "bound": "tensor" if t_tensor > t_hbm else "hbm",
                     "eff_vs_roofline": round(max(t_tensor, t_hbm) / ms, 3)})
        del x, res, x2
    out = {"batch": n, "reps": reps, "sum_ms": tot_ms, "sum_ideal_ms": sum(r["ideal_ms"] for r in rows),
           "peaks": {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops_burst": pk["bf16_tflops"]}, "layers": rows}
    print(json.dumps(out, indent=1))
    hdr = f"{'layer':24s} {'ms':>8s} {'TF/s':>7s} {'%tc':>6s} {'GB/s':>7s} {'%hbm':>6s} {'ideal':>7s} {'eff':>5s} bound"
    print(hdr, file=sys.stderr)
    for r in rows:
        print(f"{r['name']:24s} {r['ms']:8.4f} {r['tflops']:7.1f} {100*r['frac_tensor_burst']:6.1f} {r['gbs']:7.1f} "
              f"{100*r['frac_hbm']:6.1f} {r['ideal_ms']:7.4f} {r['eff_vs_roofline']:5.2f} {r['bound']}", file=sys.stderr)
    print(f"sum {tot_ms:.3f} ms, ideal {out['sum_ideal_ms']:.3f} ms", file=sys.stderr)


if __name__ == "__main__":
    main()
