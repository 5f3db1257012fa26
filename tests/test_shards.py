"""On-disk contract: what phdfx.shards writes is what the reference's reader consumes, and — given
the same clips — exactly what the reference's own writer functions produce."""
import os
import random
import sys

import pytest
import torch

from conftest import REFERENCE_SRC
from phdfx.shards import (AUG_NAMES, AsyncShardWriter, ClipRecord, ShardWriter, assemble_shard, index_from_plan,
                          plan_shards, shard_path)

HAVE_REF = os.path.isdir(REFERENCE_SRC)


def make_records(n_clips, n_vars, T=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    recs = []
    for i in range(n_clips):
        feats = [torch.randn(T, 2048, generator=g) for _ in range(n_vars)]
        j3 = [torch.randn(T, 17, 3, generator=g) * 1000 for _ in range(n_vars)]
        j2 = [torch.rand(T, 17, 2, generator=g) * 224 for _ in range(n_vars)]
        K = [torch.eye(3) * (i + 1) for _ in range(n_vars)]
        metas = [{"subject": [1, 6, 7, 8][i % 4], "action": f"Walk_{i % 3}", "cam": f"cam_{i % 4}", "start": 5 * i,
                  "end": 5 * i + T, "aug": AUG_NAMES[v] if n_vars > 1 else "orig",
                  "box": None if n_vars > 1 else torch.tensor([1, 2, 200, 200])} for v in range(n_vars)]
        recs.append(ClipRecord(feats, j3, j2, K, metas))
    return recs


def write(tmp, recs, n_vars, shard_size, pool, seed=123, fp16=False):
    w = ShardWriter(tmp, n_vars, shard_size=shard_size, shuffle_pool=pool, shuffle_seed=seed)
    for r in recs:
        w.add(r)
    return w.finish(seq_len=6, frame_skip=2, save_fp16=fp16, augment=n_vars > 1)


@pytest.mark.parametrize("n_vars", [1, 4])
def test_layout_and_index(tmp_path, n_vars):
    recs = make_records(23, n_vars)
    index = write(tmp_path, recs, n_vars, shard_size=5, pool=8)
    assert index["n_clips"] == 23 and index["n_shards"] == 5 and index["n_variants"] == n_vars
    assert index["aug_names"] == (AUG_NAMES if n_vars == 4 else ["orig"]) and index["variants_grouped"] is True
    assert sorted(os.listdir(tmp_path)) == ["index.pt"] + [f"shard_{i:05d}.pt" for i in range(5)]
    # legacy (non-zip) pickle, loadable with weights_only=True like the reference reader (dataset_features.py:107)
    with open(tmp_path / "shard_00000.pt", "rb") as f:
        assert f.read(2) != b"PK"
    idx = torch.load(tmp_path / "index.pt", map_location="cpu", weights_only=True)
    seen = set()
    for c in idx["clips"]:
        shard = torch.load(tmp_path / f"shard_{c['shard_id']:05d}.pt", map_location="cpu", weights_only=True)
        assert shard["n_vars"] == n_vars and shard["feats"].shape[1:] == (6, 2048)
        assert shard["feats"].shape[0] == shard["joints3d"].shape[0] == len(shard["meta"])
        assert c["row"] % n_vars == 0
        m = shard["meta"][c["row"]]
        assert (m["subject"], m["action"], m["cam"], m["start"]) == (c["subject"], c["action"], c["cam"], c["start"])
        # find the source record and compare every variant row
        src = next(r for r in recs if r.metas[0]["start"] == c["start"])
        for v in range(n_vars):
            assert torch.equal(shard["feats"][c["row"] + v], src.feats[v])
            assert torch.equal(shard["K"][c["row"] + v], src.K[v])
            assert shard["meta"][c["row"] + v]["aug"] == (AUG_NAMES[v] if n_vars > 1 else "orig")
        seen.add(c["start"])
    assert len(seen) == 23
    sizes = [torch.load(tmp_path / f"shard_{i:05d}.pt", weights_only=True)["feats"].shape[0] // n_vars
             for i in range(5)]
    assert sizes == [5, 5, 5, 5, 3]  # full shards + one partial (:374-396)


def test_empty_and_exact_multiple(tmp_path):
    idx = write(tmp_path / "a", [], 1, shard_size=4, pool=8)
    assert idx["n_shards"] == 0 and idx["clips"] == []
    idx = write(tmp_path / "b", make_records(8, 1), 1, shard_size=4, pool=100)
    assert idx["n_shards"] == 2


def test_writer_errors_surface(tmp_path):
    w = AsyncShardWriter()
    w.save({"x": torch.zeros(1)}, tmp_path / "no_such_dir" / "f.pt")
    with pytest.raises(RuntimeError):
        w.wait()


def test_variant_count_is_checked(tmp_path):
    w = ShardWriter(tmp_path, 4)
    with pytest.raises(ValueError):
        w.add(make_records(1, 1)[0])


@pytest.mark.skipif(not HAVE_REF, reason="reference sources only exist in the build container")
@pytest.mark.parametrize("n_vars", [1, 4])
def test_identical_to_reference_writer(tmp_path, n_vars):
    """Feed the same clips through the REFERENCE's writer functions (imported unmodified) and through ours."""
    import torchvision.io as tio

    if not hasattr(tio, "VideoReader"):
        tio.VideoReader = None  # shim: symbol removed upstream; not used by the functions called here
    sys.path.insert(0, REFERENCE_SRC)  # for the reference script's own `from dataset import ...`
    import importlib.util

    # load by path under a private name: our drop-in script has the same file name as the reference's
    spec = importlib.util.spec_from_file_location("reference_preprocess_resnet_features",
                                                  os.path.join(REFERENCE_SRC, "preprocess_resnet_features.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    recs = make_records(37, n_vars, seed=3)
    shard_size, pool, seed = 4, 10, 123
    # --- reference flow (:266-396), driven exactly like its main loop
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    writer = ref.AsyncFileWriter()
    rng = random.Random(seed)
    shuffle_pool, carry, clip_index, shard_id = [], [], [], 0
    for r in recs:
        shuffle_pool.append([{"feat": r.feats[v], "joints3d": r.joints3d[v], "joints2d": r.joints2d[v], "K": r.K[v],
                              "meta": r.metas[v]} for v in range(n_vars)])
        if len(shuffle_pool) >= pool:
            shard_id, carry = ref.flush_pool_groups_to_shards(shuffle_pool, carry, shard_id, n_vars, ref_dir, writer,
                                                              shard_size, clip_index, rng)
            shuffle_pool = []
    final = carry + shuffle_pool
    rng.shuffle(final)
    for s in range(0, len(final), shard_size):  # full shards then the partial one (:347-396)
        groups = final[s:s + shard_size]
        buf = ref.empty_shard_buffer()
        for i, g in enumerate(groups):
            m0 = g[0]["meta"]
            clip_index.append({"shard_id": shard_id, "row": i * n_vars, "subject": m0["subject"],
                               "action": m0["action"], "cam": m0["cam"], "start": m0["start"], "end": m0["end"]})
            for e in g:
                for k_src, k_dst in (("feat", "feats"), ("joints3d", "joints3d"), ("joints2d", "joints2d"), ("K", "K"),
                                     ("meta", "meta")):
                    buf[k_dst].append(e[k_src])
        ref.flush_shard(buf, shard_id, n_vars, ref_dir, writer)
        shard_id += 1
    writer.wait()
    writer.stop()
    # --- ours
    ours = write(tmp_path / "ours", recs, n_vars, shard_size, pool, seed)
    assert ours["n_shards"] == shard_id and ours["clips"] == clip_index
    for sid in range(shard_id):
        a = torch.load(ref_dir / f"shard_{sid:05d}.pt", weights_only=True)
        b = torch.load(tmp_path / "ours" / f"shard_{sid:05d}.pt", weights_only=True)
        assert a.keys() == b.keys() and a["n_vars"] == b["n_vars"]
        for k in ("feats", "joints3d", "joints2d", "K"):
            assert torch.equal(a[k], b[k]) and a[k].dtype == b[k].dtype, (sid, k)
        for ma, mb in zip(a["meta"], b["meta"]):
            assert {k: v for k, v in ma.items() if k != "box"} == {k: v for k, v in mb.items() if k != "box"}


@pytest.mark.skipif(not HAVE_REF, reason="reference sources only exist in the build container")
def test_reference_reader_and_sampler_consume_our_output(tmp_path):
    """Unmodified Human36MFeatureClips + MixedShardBatchSampler over shards written by phdfx.shards."""
    sys.path.insert(0, REFERENCE_SRC)
    from dataset_features import Human36MFeatureClips
    from samplers import MixedShardBatchSampler

    recs = make_records(32, 4, seed=5)
    write(tmp_path, recs, 4, shard_size=8, pool=16)
    ds = Human36MFeatureClips(root=str(tmp_path), subjects=[1, 6, 7, 8], augment=True, shard_cache_size=8)
    assert len(ds) == 32 * 4
    feats, j3, j2, K = ds[5]
    assert feats.shape == (6, 2048) and j3.shape == (6, 17, 3) and K.shape == (3, 3)
    # row + var_offset addressing (dataset_features.py:116): item 5 = clip 1, variant 1
    clip = ds._clips[1]
    src = next(r for r in recs if r.metas[0]["start"] == clip["start"])
    assert torch.equal(feats, src.feats[1]) and torch.allclose(j3, src.joints3d[1] / 1000.0)
    sampler = MixedShardBatchSampler(ds, batch_size=8, shards_per_batch=4, shuffle=True, seed=0)
    batches = list(sampler)
    assert batches and all(len(b) == 8 for b in batches)
    ds_test = Human36MFeatureClips(root=str(tmp_path), subjects=[6], test_set=True)
    assert all(ds_test[i][4]["subject"] == 6 for i in range(len(ds_test)))


# ---------------------------------------------------------------------------------------------- shard plan
@pytest.mark.parametrize("n_clips,shard_size,pool", [(0, 4, 8), (1, 4, 8), (3, 4, 8), (8, 4, 8), (23, 5, 8), (23, 5, 7),
                                                      (40, 4, 6), (64, 8, 8), (100, 7, 1000), (57, 1, 3)])
def test_plan_shards_is_the_streaming_writers_permutation(tmp_path, n_clips, shard_size, pool):
    """plan_shards (integers only) predicts exactly which clip the streaming writer — itself byte-identical to the
    reference's functions (test_identical_to_reference_writer) — puts in which row of which shard."""
    recs = make_records(n_clips, 1)
    for i, r in enumerate(recs):
        r.metas[0]["start"] = i  # tag every record with its arrival number
    index = write(tmp_path, recs, 1, shard_size=shard_size, pool=pool, seed=77)
    plan = plan_shards(n_clips, shard_size, pool, 77)
    assert len(plan) == index["n_shards"] and sum(len(p) for p in plan) == n_clips
    assert sorted(i for p in plan for i in p) == list(range(n_clips))
    for sid, ids in enumerate(plan):
        shard = torch.load(tmp_path / f"shard_{sid:05d}.pt", weights_only=True)
        assert [m["start"] for m in shard["meta"]] == ids
    assert [(c["shard_id"], c["row"], c["start"]) for c in index["clips"]] == \
        [(sid, row, i) for sid, ids in enumerate(plan) for row, i in enumerate(ids)]


@pytest.mark.parametrize("n_vars", [1, 4])
def test_planned_writing_equals_streaming(tmp_path, n_vars):
    """Shards assembled rank by rank from the plan (any order, any owner) + index_from_plan == what the one-process
    streaming writer puts on disk: same index, same tensors bit for bit, same meta lists, same pickle format.  (The
    raw bytes of legacy torch pickles embed storage keys derived from memory addresses, so they are not comparable.)"""
    recs = make_records(29, n_vars, seed=3)
    a, b = tmp_path / "stream", tmp_path / "planned"
    index = write(a, recs, n_vars, shard_size=6, pool=10, seed=5)
    b.mkdir()
    plan = plan_shards(29, 6, 10, 5)
    w = AsyncShardWriter()
    for rank in (1, 0, 2):  # three "ranks", out of order
        for sid in range(rank, len(plan), 3):
            w.save(assemble_shard([recs[i] for i in plan[sid]], n_vars), shard_path(b, sid))
    w.wait()
    w.stop()
    idx = index_from_plan(plan, lambda i: recs[i].metas[0], n_vars, seq_len=6, frame_skip=2, save_fp16=False,
                          augment=n_vars > 1, shuffle_seed=5, shuffle_pool=10)
    assert idx == index
    for sid in range(len(plan)):
        x = torch.load(a / f"shard_{sid:05d}.pt", weights_only=True)
        y = torch.load(b / f"shard_{sid:05d}.pt", weights_only=True)
        assert x.keys() == y.keys() and x["n_vars"] == y["n_vars"] == n_vars
        for k in ("feats", "joints3d", "joints2d", "K"):
            assert x[k].dtype == y[k].dtype and torch.equal(x[k], y[k]), k
        assert len(x["meta"]) == len(y["meta"])
        for mx, my in zip(x["meta"], y["meta"]):
            assert {k: v for k, v in mx.items() if k != "box"} == {k: v for k, v in my.items() if k != "box"}
            assert (mx["box"] is None and my["box"] is None) or torch.equal(mx["box"], my["box"])
        # storage keys are decimal renderings of addresses: the sizes may differ by a few digits, not more
        assert abs((a / f"shard_{sid:05d}.pt").stat().st_size - (b / f"shard_{sid:05d}.pt").stat().st_size) <= 64
        with open(b / f"shard_{sid:05d}.pt", "rb") as f:
            assert f.read(2) != b"PK"  # legacy (non-zip) pickle, like the reference (:45)


def test_plan_shards_property_based(tmp_path_factory):
    """Random (n_clips, shard_size, shuffle_pool, seed): the integer-only plan always equals what the streaming writer
    does to tagged records (tiny tensors, so hundreds of cases stay cheap)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(0, 60), shard=st.integers(1, 9), pool=st.integers(1, 25), seed=st.integers(0, 2 ** 31 - 1))
    def check(n, shard, pool, seed):
        root = tmp_path_factory.mktemp("plan")
        w = ShardWriter(root, 1, shard_size=shard, shuffle_pool=pool, shuffle_seed=seed)
        for i in range(n):
            meta = {"subject": 1, "action": "a", "cam": "cam_0", "start": i, "end": i + 1, "aug": "orig", "box": None}
            w.add(ClipRecord([torch.full((1, 4), float(i))], [torch.zeros(1, 17, 3)], [torch.zeros(1, 17, 2)],
                             [torch.eye(3)], [meta]))
        index = w.finish(seq_len=1, frame_skip=1, save_fp16=False, augment=False)
        plan = plan_shards(n, shard, pool, seed)
        assert [[c["start"] for c in index["clips"] if c["shard_id"] == s] for s in range(index["n_shards"])] == plan
        assert all(len(p) == shard for p in plan[:-1]) and (not plan or 1 <= len(plan[-1]) <= shard)
        for sid, ids in enumerate(plan):
            feats = torch.load(root / f"shard_{sid:05d}.pt", weights_only=True)["feats"]
            assert feats[:, 0, 0].tolist() == [float(i) for i in ids]

    check()
