"""The "second baseline" of BASELINE.md §4: the reference's own GPU path on the same B200 — torchvision ResNet-50 minus
the fc layer, eval mode, eager cuDNN under torch.autocast(bf16) with cudnn.benchmark and TF32 allowed, fp32 NCHW
normalised input (reference src/preprocess_resnet_features.py:164-167, 207-209, 288-297), optionally wrapped in
torch.compile(mode="max-autotune") as the reference does on one GPU (:220-226).  A comparison arm only: nothing in the
product path, the tests or bench.py uses it.

    python tools/bench_cudnn_arm.py eager [batch]      # one JSON line
    python tools/bench_cudnn_arm.py compile [batch]    # may take minutes (autotuning); run under `timeout`

Two figures per mode: `device_resident` (input already in HBM; what bench.py's `value` is for our path) and `host_loop`
(the reference's per-batch loop: pinned fp32 clip `.to(device, non_blocking=True)`, forward, `.float().cpu()`; what
bench.py's `e2e` is for our path — the reference moves fp32 NCHW, 602 KB per frame, where Seam B moves uint8 HWC, 150 KB).
"""
import json
import sys
import time

import torch
import torch.nn as nn
from torchvision import models


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "eager"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(0)
    resnet = models.resnet50(weights=None)  # no network: random init, like every BASELINE config
    backbone = nn.Sequential(*list(resnet.children())[:-1]).to("cuda").eval()
    t0 = time.time()
    if mode == "compile":
        backbone = torch.compile(backbone, mode="max-autotune")
    T = 8
    g = torch.Generator().manual_seed(2)
    host = [torch.randn(n // T, T, 3, 224, 224, generator=g).pin_memory() for _ in range(2)]
    dev = [h.cuda() for h in host]

    def fwd(v_video):
        Bv, Tt, C, H, W = v_video.shape
        with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            x = v_video.view(Bv * Tt, C, H, W).contiguous()
            return backbone(x).flatten(1).view(Bv, Tt, -1)

    for i in range(5):  # cudnn.benchmark / autotune / compile happen here
        fwd(dev[i & 1])
    torch.cuda.synchronize()
    setup_s = time.time() - t0

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_dev = timed(lambda i: fwd(dev[i & 1]), 20)

    def host_step(i):
        return fwd(host[i & 1].to("cuda", non_blocking=True)).to(torch.float32).cpu()

    host_step(0)
    ms_host = timed(host_step, 10)
    print(json.dumps({
        "impl": "reference GPU path (torchvision + cuDNN, autocast bf16, cudnn.benchmark)" +
                (" + torch.compile(max-autotune)" if mode == "compile" else ""),
        "batch": n, "device_resident": {"ms_per_batch": ms_dev, "frames_per_s": n / ms_dev * 1e3},
        "host_loop": {"ms_per_batch": ms_host, "frames_per_s": n / ms_host * 1e3,
                      "h2d_bytes_per_batch": n * 3 * 224 * 224 * 4, "d2h_bytes_per_batch": n * 2048 * 4},
        "setup_s": round(setup_s, 1), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
        "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
