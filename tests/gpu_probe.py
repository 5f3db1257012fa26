"""GPU diagnostic (not a pytest file): per-layer and whole-trunk comparison of libphdfx against fp32 PyTorch on
the SAME bf16-rounded operands.  Writes gpurun_out/probe_<group>.json.  Groups run in separate processes so a
fault in one kernel family does not hide the others:

    python tests/gpu_probe.py tiled|im2col|stem|gap|pool|pre|full [n_frames]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "implementation-phd-lab-vision_b200"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import torchvision  # noqa: E402

import phdfx  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def build(seed=0):
    torch.manual_seed(seed)
    r = torchvision.models.resnet50(weights=None)
    bb = torch.nn.Sequential(*list(r.children())[:-1]).eval()
    phdfx.randomize_bn_(bb, 1)
    return bb


def layer_modules(bb):
    """name -> (conv, bn) in plan order."""
    out = {"conv1": (bb[0], bb[1])}
    for li in range(4):
        for bi, blk in enumerate(bb[4 + li]):
            p = f"layer{li + 1}.{bi}"
            out[p + ".conv1"] = (blk.conv1, blk.bn1)
            out[p + ".conv2"] = (blk.conv2, blk.bn2)
            out[p + ".conv3"] = (blk.conv3, blk.bn3)
            if blk.downsample is not None:
                out[p + ".downsample"] = (blk.downsample[0], blk.downsample[1])
    return out


def to_nhwc4p(x_nchw):
    """fp32 [N,3,224,224] -> bf16 NHWC4p [N,224,232,4] (torch restatement of the layout, for probes)."""
    n = x_nchw.shape[0]
    out = torch.zeros(n, 224, 232, 4, device=x_nchw.device, dtype=torch.bfloat16)
    out[:, :, 4:228, 0:3] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out


def ref_layer(L, conv, bn, x_nhwc_bf16, res_nhwc_bf16):
    """fp32 reference on bf16-rounded operands: conv(w') + b' (+res) (+relu); NHWC fp32 out."""
    w, b = phdfx.fold_conv_bn(conv, bn)
    w = w.to(torch.bfloat16).to(torch.float32).cuda()
    b = b.cuda()
    x = x_nhwc_bf16.to(torch.float32).permute(0, 3, 1, 2).contiguous()
    y = F.conv2d(x, w, b, stride=conv.stride, padding=conv.padding)
    if res_nhwc_bf16 is not None:
        y = y + res_nhwc_bf16.to(torch.float32).permute(0, 3, 1, 2)
    if L.relu:
        y = torch.relu(y)
    if L.gap:
        return y.mean(dim=(2, 3))
    return y.permute(0, 2, 3, 1).contiguous()


def err_stats(got, ref):
    got = got.to(torch.float32)
    ref = ref.to(torch.float32)
    d = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-12
    return {
        "max_abs": d.max().item(),
        "ref_max": scale,
        "norm_err": d.max().item() / scale,
        "mean_abs": d.mean().item(),
        "nan": bool(torch.isnan(got).any().item()),
        "frac_bad": (d > 0.02 * scale).float().mean().item(),
    }


def main():
    group = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(ROOT, "gpurun_out", f"probe_{group}.json")
    results = {"group": group, "n": n, "cases": []}

    def flush():
        with open(out_path, "w") as f:
            json.dump(results, f, indent=1)

    flush()
    bb = build()
    t0 = time.time()
    eng = phdfx.B200Backbone(bb, device=0, max_frames=max(n, 8))
    results["create_s"] = time.time() - t0
    mods = layer_modules(bb)
    g = torch.Generator(device="cuda").manual_seed(7)

    def want(i, L):
        if L.kind == 2:
            return group == "pool"
        if L.kind in (1, 3):
            return group == "stem"
        if L.gap:
            return group == "gap"
        tiled = L.r == 1 and L.stride == 1
        return group == ("tiled" if tiled else "im2col")

    if group in ("tiled", "im2col", "stem", "gap", "pool"):
        seen = set()
        for i, (name, L) in enumerate(zip(eng.plan.names, eng.plan.layers)):
            if not want(i, L):
                continue
            key = (L.kind, L.cin, L.cout, L.r, L.stride, L.hin, L.res_buf >= 0, L.relu, L.gap)
            if key in seen:
                continue
            seen.add(key)
            case = {"layer": i, "name": name, **L.as_dict()}
            try:
                if L.kind in (1, 3):
                    x = torch.randn(n, 3, 224, 224, device="cuda", generator=g)
                    xin = to_nhwc4p(x)
                    x_ref = x.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
                else:
                    xin = (torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g)).to(torch.bfloat16)
                    x_ref = xin
                if L.kind == 2:
                    xin = torch.relu(xin)
                    got = eng.run_layer(i, xin)
                    ref = F.max_pool2d(xin.to(torch.float32).permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
                else:
                    ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
                    res = None
                    if L.res_buf >= 0:
                        res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16)
                    got = eng.run_layer(i, xin, res)
                    conv, bn = mods[name]
                    ref = ref_layer(L, conv, bn, x_ref, res)
                torch.cuda.synchronize()
                case.update(err_stats(got, ref))
                case["ok"] = (not case["nan"]) and case["norm_err"] < 2e-2
                if not case["ok"] and got.dim() == 4:
                    # where are the errors?  per-frame / per-row-block / per-channel-block summary
                    d = (got.to(torch.float32) - ref).abs()
                    case["err_by_frame"] = d.amax(dim=(1, 2, 3)).tolist()
                    case["err_by_row"] = d.amax(dim=(0, 2, 3)).tolist()[:16]
                    case["err_by_col"] = d.amax(dim=(0, 1, 3)).tolist()[:16]
                    cb = d.amax(dim=(0, 1, 2))
                    case["err_by_ch8"] = cb.view(-1, 8).amax(dim=1).tolist()[:32]
                    case["got_sample"] = got[0, 0, 0, :8].to(torch.float32).tolist()
                    case["ref_sample"] = ref[0, 0, 0, :8].tolist()
            except Exception as e:  # noqa: BLE001
                case["ok"] = False
                case["exception"] = repr(e)
                results["cases"].append(case)
                flush()
                print(json.dumps(case))
                break
            results["cases"].append(case)
            flush()
            print(json.dumps({k: case[k] for k in ("name", "ok", "norm_err", "max_abs", "ref_max", "frac_bad")}))
    elif group == "pre":
        import torchvision.transforms.functional as TF

        gen = torch.Generator().manual_seed(1)
        for (H, W, box) in [(300, 280, (20, 30, 217, 217)), (224, 224, None), (480, 500, (0, 100, 400, 400))]:
            frames = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, generator=gen)
            top, left, hh, ww = box if box else (0, 0, H, W)
            crop = frames.permute(0, 3, 1, 2)[:, :, top:top + hh, left:left + ww]
            r = TF.resize(crop, [224, 224], antialias=False).to(torch.float32) / 255.0
            mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
            ref = ((r - mean) / std).to(torch.bfloat16)
            boxes = None
            if box:
                boxes = torch.tensor([box] * n, dtype=torch.int32, device="cuda")
            got = eng.preprocess_u8(frames.cuda(), boxes)
            torch.cuda.synchronize()
            got_img = got[:, :, 4:228, 0:3].permute(0, 3, 1, 2).cpu()
            pad_zero = bool((got[:, :, :4].abs().sum() + got[:, :, 228:].abs().sum() + got[..., 3].abs().sum()).item() == 0)
            mism = (got_img.view(torch.int16) != ref.view(torch.int16)).float().mean().item()
            case = {"H": H, "W": W, "box": box, "mismatch_frac": mism, "pad_zero": pad_zero,
                    "max_abs": (got_img.float() - ref.float()).abs().max().item(), "ok": mism < 1e-3 and pad_zero}
            results["cases"].append(case)
            print(json.dumps(case))
            flush()
    elif group == "full":
        x = torch.randn(n, 3, 224, 224, generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            ref = bb(x).flatten(1)
        got = eng(x.cuda()).flatten(1).cpu()
        torch.cuda.synchronize()
        d = (got - ref).abs()
        per_frame = (d.amax(dim=1) / ref.abs().amax(dim=1)).tolist()
        cos = F.cosine_similarity(got, ref, dim=1).tolist()
        case = {"norm_err_per_frame": per_frame, "cos_per_frame": cos, "launches": eng.launches,
                "ok": max(per_frame) <= 2e-2 and min(cos) >= 0.9999, "nan": bool(torch.isnan(got).any())}
        results["cases"].append(case)
        print(json.dumps(case))
        # timing, rough
        xs = x.cuda()
        for _ in range(2):
            eng(xs)
        torch.cuda.synchronize()
        t = time.time()
        for _ in range(5):
            eng(xs)
        torch.cuda.synchronize()
        results["ms_per_call"] = (time.time() - t) / 5 * 1e3
        print("ms per call", results["ms_per_call"], "n", n)
    results["all_ok"] = all(c.get("ok") for c in results["cases"]) and len(results["cases"]) > 0
    flush()
    print("ALL_OK" if results["all_ok"] else "SOME_FAILED", group)


if __name__ == "__main__":
    main()
