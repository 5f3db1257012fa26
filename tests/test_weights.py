"""Host logic: BN folding, weight packing, execution list (phdfx/weights.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import phdfx
import resnet50_ref as R


@pytest.fixture(scope="module")
def backbone():
    return R.seeded_backbone()


@pytest.fixture(scope="module")
def plan(backbone):
    return phdfx.build_plan(backbone)


def test_fold_is_exact_in_fp32(backbone):
    blk = backbone[5][0]  # layer2.0 (has stride + downsample)
    x = torch.randn(2, blk.conv2.in_channels, 12, 12)
    for conv, bn in ((blk.conv2, blk.bn2), (blk.downsample[0], blk.downsample[1])):
        xin = x if conv is blk.conv2 else torch.randn(2, conv.in_channels, 12, 12)
        with torch.no_grad():
            ref = bn(conv(xin))
            w, b = phdfx.fold_conv_bn(conv, bn)
            got = F.conv2d(xin, w, b, stride=conv.stride, padding=conv.padding)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)


def test_plan_structure(plan, backbone):
    kinds = [l.kind for l in plan.layers]
    assert kinds[0] == 3 and kinds.count(0) == 48  # fused stem+maxpool, 48 launches for the 52 bottleneck convs
    assert len(plan.layers) == 49 and plan.names[-1] == "layer4.2.conv3"
    assert sum(1 for l in plan.layers if l.in2_buf >= 0) == 4  # down-sample branches folded into conv3
    unfused = phdfx.build_plan(backbone, fuse_stem_pool=False, fuse_downsample=False)
    assert [l.kind for l in unfused.layers[:2]] == [1, 2] and len(unfused.layers) == 54
    assert sum(1 for l in unfused.layers if l.res_buf >= 0) == 16
    assert sum(1 for n, l in zip(unfused.names, unfused.layers) if n.endswith("downsample") and l.relu == 0) == 4
    assert sum(l.gap for l in plan.layers) == 1 and plan.layers[-1].gap == 1
    # 12 conv3 with an identity residual; the 4 first-block conv3 carry the down-sample branch as a second source
    assert sum(1 for l in plan.layers if l.res_buf >= 0) == 12
    # MAC count per frame: 4 087 136 256 including the stem
    macs = 0
    for l in plan.layers:
        if l.kind == 2:
            continue
        ho = (l.hin + 2 * l.pad - l.r) // l.stride + 1
        macs += ho * ho * l.cout * (l.cin * l.r * l.s + (l.cin2 if l.in2_buf >= 0 else 0))
    assert macs == 4_087_136_256
    # no layer writes a buffer it reads
    for l in plan.layers:
        assert l.out_buf != l.in_buf and (l.res_buf < 0 or l.res_buf != l.out_buf or l.gap)
    # offsets are 128-byte aligned and inside the blobs
    for l in plan.layers:
        if l.kind != 2:
            assert l.w_off % 64 == 0 and l.w_off < plan.weights.numel() and l.b_off + l.cout <= plan.bias.numel()


def test_dataflow_is_consistent(plan):
    """Every buffer a layer reads was last written with the shape it expects."""
    shape = {0: (224, 3)}
    for name, l in zip(plan.names, plan.layers):
        h, c = shape[l.in_buf]
        assert (h, c) == (l.hin, l.cin), name
        ho = (l.hin + 2 * l.pad - l.r) // l.stride + 1
        if l.kind == 3:
            ho = 56
        if l.res_buf >= 0:
            assert shape[l.res_buf] == (ho, l.cout), name
        if l.in2_buf >= 0:
            assert shape[l.in2_buf] == (l.hin2, l.cin2) and (l.hin2 - 1) // l.stride2 + 1 == ho, name
        shape[l.out_buf] = (ho, l.cout)


def test_pack_conv_layout():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)  # [Cout,Cin,R,S]
    p = phdfx.pack_conv(w).to(torch.float32).reshape(2, 3, 3, 3)  # [Cout,R,S,Cin]
    for co in range(2):
        for r in range(3):
            for s in range(3):
                for ci in range(3):
                    assert p[co, r, s, ci] == w[co, ci, r, s]


def test_pack_stem_layout():
    w = torch.randn(64, 3, 7, 7)
    p = phdfx.pack_stem(w).to(torch.float32).reshape(7, 64, 8, 4)
    assert torch.count_nonzero(p[:, :, 0]) == 0 and torch.count_nonzero(p[..., 3]) == 0
    wb = w.to(torch.bfloat16).to(torch.float32)
    for r in (0, 3, 6):
        for s in (0, 2, 6):
            assert torch.equal(p[r, :, s + 1, :3], wb[:, :, r, s])


def test_pack_stem_pool_layout():
    w = torch.randn(64, 3, 7, 7)
    full = phdfx.pack_stem_pool(w).to(torch.float32)
    assert full.numel() == 7 * 4 * 64 * 8 + 5 * 4 * 128 * 8
    p = full[:7 * 4 * 64 * 8].reshape(7, 4, 64, 8)  # [r][k-chunk][cout][e], k = chunk*8 + e
    pairs = full[7 * 4 * 64 * 8:].reshape(5, 4, 128, 8)  # stacked rows for the N = 128 MMAs: [e-2][chunk][e | e-2][8]
    for e in range(2, 7):
        assert torch.equal(pairs[e - 2, :, :64], p[e]) and torch.equal(pairs[e - 2, :, 64:], p[e - 2])
    wb = w.to(torch.bfloat16).to(torch.float32)
    for r in (0, 2, 6):
        for s in (0, 3, 6):
            for c in range(3):
                k = (s + 1) * 4 + c
                assert torch.equal(p[r, k // 8, :, k % 8], wb[:, c, r, s])
    # tap -1 (k = 0..3) and the padding channel carry zero weights
    assert torch.count_nonzero(p[:, 0, :, 0:4]) == 0
    assert torch.count_nonzero(p.reshape(7, 4, 64, 2, 4)[..., 3]) == 0


def test_packed_weights_reproduce_the_network(backbone, plan):
    """Unpack layer2.0.conv2 from the blob and check it against the module (closes the loop pack -> offsets)."""
    # fused conv3 + down-sample: weights [cout][width | cin], bias b3 + bd
    j = plan.names.index("layer2.0.conv3+downsample")
    lf = plan.layers[j]
    kf = lf.cin + lf.cin2
    wf = plan.weights[lf.w_off:lf.w_off + lf.cout * kf].to(torch.float32).reshape(lf.cout, kf)
    blk0 = backbone[5][0]
    w3, b3 = phdfx.fold_conv_bn(blk0.conv3, blk0.bn3)
    wd, bd = phdfx.fold_conv_bn(blk0.downsample[0], blk0.downsample[1])
    assert torch.equal(wf[:, :lf.cin], w3.reshape(lf.cout, -1).to(torch.bfloat16).to(torch.float32))
    assert torch.equal(wf[:, lf.cin:], wd.reshape(lf.cout, -1).to(torch.bfloat16).to(torch.float32))
    assert torch.equal(plan.bias[lf.b_off:lf.b_off + lf.cout], b3 + bd)
    i = plan.names.index("layer2.0.conv2")
    l = plan.layers[i]
    k = l.r * l.s * l.cin
    w = plan.weights[l.w_off:l.w_off + l.cout * k].to(torch.float32).reshape(l.cout, l.r, l.s, l.cin)
    blk = backbone[5][0]
    wf, bf = phdfx.fold_conv_bn(blk.conv2, blk.bn2)
    assert torch.equal(w.permute(0, 3, 1, 2), wf.to(torch.bfloat16).to(torch.float32))
    assert torch.equal(plan.bias[l.b_off:l.b_off + l.cout], bf)


def test_randomize_bn_matches_oracle_builder(backbone):
    import torchvision

    torch.manual_seed(R.WEIGHT_SEED)
    bb = torch.nn.Sequential(*list(torchvision.models.resnet50(weights=None).children())[:-1]).eval()
    phdfx.randomize_bn_(bb, R.BN_SEED)
    for (n1, p1), (n2, p2) in zip(bb.state_dict().items(), backbone.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2), n1


def test_jitter_params_row():
    """phdfx.jitter_params: K1's per-frame colour-jitter row from torchvision's ColorJitter.make_params dict."""
    import numpy as np
    from torchvision.transforms import v2 as T2

    torch.manual_seed(3)
    prm = T2.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.2, hue=0.05).make_params([])
    row = phdfx.jitter_params(**prm)
    assert row.dtype == torch.float32 and tuple(row.shape) == (12,)
    assert [int(v) for v in row[:4]] == [int(v) for v in prm["fn_idx"]]
    c, s = float(prm["contrast_factor"]), float(prm["saturation_factor"])
    want = [float(prm["brightness_factor"]), c, 1.0 - c, s, 1.0 - s, float(prm["hue_factor"])]
    assert np.array_equal(row[4:10].numpy(), np.array(want, dtype=np.float64).astype(np.float32))
    assert torch.equal(row, phdfx.jitter_params(prm["fn_idx"], prm["brightness_factor"], c, s, prm["hue_factor"]))
    with pytest.raises(ValueError):
        phdfx.jitter_params([0, 1, 2, 2], 1.0, 1.0, 1.0, 0.0)
    with pytest.raises(ValueError):
        phdfx.jitter_params([0, 1, 2, 3], 1.0, 1.0, 1.0)


def test_product_and_oracle_build_the_same_seeded_data():
    """bench.py and `--synthetic` runs take weights / frames from phdfx.synthetic; the oracle keeps its own copy of the
    recipe (nothing under phdfx/ may import oracle/).  The two must stay the same tensors and bytes."""
    import numpy as np
    import resnet50_ref as R
    from phdfx.synthetic import csrc_sha, seeded_backbone, seeded_frames

    a, b = seeded_backbone().state_dict(), R.seeded_backbone().state_dict()
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert np.array_equal(seeded_frames(3, 17, 19, 5), R.seeded_frames(3, 17, 19, 5))
    assert len(csrc_sha()) == 16
