"""BASELINE config 5, GPU half: extract features for a fixed synthetic clip set with the B200 path and save them.
The CPU half (tests/test_e2e_loss.py, run where the reference sources exist) feeds these features and the reference's
own fp32 features to the UNMODIFIED reference reader / sampler / model / train step and compares the losses.

    python tools/e2e_extract.py gpurun_out/e2e_b200_feats.npz      # on the GPU box
    cp gpurun_out/e2e_b200_feats.npz tests/golden/                   # then commit the fixture

The fixture carries the content hash of the kernel sources it was extracted with (phdfx.csrc_sha);
tests/test_gpu_parity.py::test_config5_fixture_is_what_head_extracts re-extracts the same clips on the GPU box and
compares, so the CPU-side loss test is tied to the kernels that ship.
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "implementation-phd-lab-vision_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

E2E = dict(n_clips=16, seq_len=16, height=260, width=300, box_side=241, subjects=(1, 6, 7, 8), seed=0)


def main():
    import phdfx
    from phdfx.synthetic import SyntheticH36MClips, csrc_sha, seeded_backbone

    out = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "gpurun_out" / "e2e_b200_feats.npz")
    ds = SyntheticH36MClips(**E2E)
    eng = phdfx.B200Backbone(seeded_backbone(), device=0, max_frames=E2E["n_clips"] * E2E["seq_len"])
    frames = torch.stack([ds.frames(i) for i in range(len(ds))])  # (N,T,H,W,3)
    boxes = torch.stack([ds.box(i) for i in range(len(ds))]).to(torch.int32)
    N, T = frames.shape[:2]
    fr = frames.view(N * T, *frames.shape[2:]).cuda()
    bx = boxes.repeat_interleave(T, dim=0).cuda()
    feats = eng.extract_u8(fr, bx).view(N, T, 2048).cpu().numpy()
    np.savez_compressed(out, feats=feats.astype(np.float32), cfg=np.array(
        [E2E["n_clips"], E2E["seq_len"], E2E["height"], E2E["width"], E2E["box_side"], E2E["seed"]], dtype=np.int64),
        device=np.array(torch.cuda.get_device_name(0)), csrc_sha=np.array(csrc_sha()))
    print("wrote", out, feats.shape, "launches", eng.launches)


if __name__ == "__main__":
    main()
