"""BASELINE config 5: features extracted by the B200 path (tests/golden/e2e_b200_feats.npz, produced on a B200 by
tools/e2e_extract.py) vs the reference's own CPU fp32 features, both fed to the UNMODIFIED reference
dataset_features.Human36MFeatureClips -> samplers.MixedShardBatchSampler -> model.PHDFor3DJoints(1024, 2) ->
train.train() for one epoch with identical seeds.  Pass: |loss_b200 - loss_ref| <= 1e-3 * max(1, |loss_ref|).

CPU test, build container only (needs /root/reference); the GPU-side numbers come from the committed fixture."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE_SRC

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "e2e_b200_feats.npz")
pytestmark = pytest.mark.skipif(not (os.path.isdir(REFERENCE_SRC) and os.path.exists(FIXTURE)),
                                reason="needs the reference sources and the GPU-extracted fixture")


def _write_root(root, feats, ds):
    from phdfx.shards import ClipRecord, ShardWriter

    w = ShardWriter(root, 1, shard_size=4, shuffle_pool=8, shuffle_seed=123)
    for i in range(len(ds)):
        j3, j2, K = ds.annotations(i)
        c = ds.index[i]
        meta = {"subject": c.subject, "action": c.action, "cam": c.cam, "start": c.start, "end": c.end, "aug": "orig",
                "box": ds.box(i)}
        w.add(ClipRecord([torch.from_numpy(feats[i])], [j3], [j2], [K], [meta]))
    return w.finish(seq_len=ds.seq_len, frame_skip=2, save_fp16=False, augment=False)


def _one_epoch(root):
    sys.path.insert(0, REFERENCE_SRC)
    from dataset_features import Human36MFeatureClips  # the reference, unmodified
    from model import PHDFor3DJoints as PHD
    from samplers import MixedShardBatchSampler
    from train import train

    torch.manual_seed(0)
    ds = Human36MFeatureClips(root=root, subjects=[1, 6, 7, 8], augment=False, shard_cache_size=8)
    sampler = MixedShardBatchSampler(ds, batch_size=8, shards_per_batch=4, shuffle=True, seed=0)
    loader = torch.utils.data.DataLoader(ds, batch_sampler=sampler, num_workers=0)
    model = PHD(latent_dim=1024, number_blocks=2)  # train.py:370
    for p in model.f_AR.parameters():  # train.py:375-376
        p.requires_grad = False
    optim = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4)
    torch.manual_seed(1)  # dropout stream
    loss, mpjpe = train(model, loader, optim, scaler=None, device=torch.device("cpu"), log_every=0)
    return loss, mpjpe


def test_downstream_loss_parity(tmp_path):
    import preprocess_ref as P
    import resnet50_ref as R
    from phdfx.synthetic import SyntheticH36MClips

    fx = np.load(FIXTURE)
    n_clips, seq_len, H, W, side, seed = (int(v) for v in fx["cfg"])
    ds = SyntheticH36MClips(n_clips, seq_len=seq_len, height=H, width=W, subjects=(1, 6, 7, 8), seed=seed,
                            box_side=side)
    b200 = fx["feats"]
    assert b200.shape == (n_clips, seq_len, 2048)
    # reference features: the reference's front end + torchvision fp32 trunk (:207-209, :296)
    bb = R.seeded_backbone()
    ref = np.empty_like(b200)
    with torch.no_grad():
        for i in range(n_clips):
            x = P.crop_resize_normalize(ds.frames(i).numpy(), ds.box(i).tolist())
            ref[i] = bb(torch.from_numpy(x)).flatten(1).numpy()
    err = np.abs(b200 - ref).max(axis=2) / np.abs(ref).max(axis=2)
    cos = (b200 * ref).sum(2) / (np.linalg.norm(b200, axis=2) * np.linalg.norm(ref, axis=2))
    assert err.max() <= 2e-2 and cos.min() >= 0.9999, (err.max(), cos.min())
    _write_root(tmp_path / "ref", ref, ds)
    _write_root(tmp_path / "b200", b200, ds)
    loss_ref, mpjpe_ref = _one_epoch(str(tmp_path / "ref"))
    loss_b200, mpjpe_b200 = _one_epoch(str(tmp_path / "b200"))
    print(f"loss ref {loss_ref:.6f}  b200 {loss_b200:.6f}  |d| {abs(loss_ref - loss_b200):.2e}; "
          f"mpjpe ref {mpjpe_ref:.4f} b200 {mpjpe_b200:.4f}; feat norm_err {err.max():.2e} min cos {cos.min():.6f}")
    assert abs(loss_b200 - loss_ref) <= 1e-3 * max(1.0, abs(loss_ref))
