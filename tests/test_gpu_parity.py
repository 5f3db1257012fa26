"""Parity tests proper (run on the B200 box with -m gpu).  Everything goes through the C ABI (libphdfx.so via
phdfx.B200Backbone); the checker is the oracle (oracle/*.py, oracle/resnet50_ref.c) and the committed golden vectors
generated from the reference's own arithmetic.  Nothing here reads /root/reference.

Tolerance (BASELINE.json north_star): per frame  max|got - ref| / max|ref| <= 2e-2  and  cosine >= 0.9999  for the
bf16 trunk against the fp32 reference; bit-exact for the uint8 resize and the bf16 NHWC4p bytes K1 writes.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import phdfx
import preprocess_ref as P
import resnet50_ref as R

pytestmark = pytest.mark.gpu

NORM_TOL = 2e-2
COS_TOL = 0.9999


def frame_errors(got: np.ndarray, ref: np.ndarray):
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    return err, cos


@pytest.fixture(scope="module")
def backbone():
    return R.seeded_backbone()


@pytest.fixture(scope="module")
def eng(backbone):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    e = phdfx.B200Backbone(backbone, device=0, max_frames=32)
    yield e
    e.close()


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "trunk_golden.npz"))


# ------------------------------------------------------------------------------------------------ K1 preprocess
def _pre_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess_golden.npz"))
    i = 0
    while f"case{i}_meta" in g:
        n, H, W, top, left, hh, ww, seed = (int(v) for v in g[f"case{i}_meta"])
        yield g, i, R.seeded_frames(n, H, W, seed), (top, left, hh, ww)
        i += 1


def test_preprocess_bit_exact_vs_reference_golden(eng, golden_dir):
    """K1 output bytes == bf16(normalise(reference resize golden)) — bit-exact, incl. padding columns/channel."""
    for g, i, frames, box in _pre_cases(golden_dir):
        n = frames.shape[0]
        boxes = torch.tensor([box] * n, dtype=torch.int32, device="cuda")
        got = eng.preprocess_u8(torch.from_numpy(frames).cuda(), boxes)
        u8 = g[f"case{i}_u8_worker"].astype(np.float32)
        x = (u8 / np.float32(255.0) - P.IMAGENET_MEAN[None, :, None, None]) / P.IMAGENET_STD[None, :, None, None]
        want = P.to_nhwc4p_bf16_bits(x.astype(np.float32))
        got_bits = got.view(torch.int16).cpu().numpy().view(np.uint16)
        assert np.array_equal(got_bits, want), f"case {i} box {box}: {(got_bits != want).sum()} differing values"


def test_preprocess_matches_oracle_on_ragged_boxes(eng):
    """Per-frame boxes of different sizes/positions in one call (the reference uses one box per clip)."""
    frames = R.seeded_frames(5, 333, 417, 21)
    boxes = [(0, 0, 333, 333), (10, 100, 224, 224), (50, 60, 97, 97), (332, 416, 1, 1), (3, 5, 301, 301)]
    got = eng.preprocess_u8(torch.from_numpy(frames).cuda(), torch.tensor(boxes, dtype=torch.int32, device="cuda"))
    got_bits = got.view(torch.int16).cpu().numpy().view(np.uint16)
    for k, box in enumerate(boxes):
        want = P.to_nhwc4p_bf16_bits(P.crop_resize_normalize(frames[k:k + 1], box))
        assert np.array_equal(got_bits[k:k + 1], want), f"frame {k} box {box}"


def test_preprocess_hflip(eng):
    """flip_w mirrors the resized clip (src/dataset.py:166), bit-exact."""
    frames = R.seeded_frames(2, 260, 300, 22)
    box = (7, 9, 240, 240)
    boxes = torch.tensor([box] * 2, dtype=torch.int32, device="cuda")
    got = eng.preprocess_u8(torch.from_numpy(frames).cuda(), boxes, flip_w=True)
    want = P.to_nhwc4p_bf16_bits(P.hflip(P.crop_resize_normalize(frames, box)))
    assert np.array_equal(got.view(torch.int16).cpu().numpy().view(np.uint16), want)


def _jitter_rows(cases, frames_per_case):
    rows = [phdfx.jitter_params(*c) for c in cases for _ in range(frames_per_case)]
    return torch.stack(rows).cuda()


@pytest.mark.parametrize("H,W,box", [(224, 224, (0, 0, 224, 224)), (300, 280, (10, 20, 231, 231))])
def test_preprocess_color_jitter_vs_oracle(eng, H, W, box):
    """K1's colour-jitter variant (the reference's `cjitter`, src/dataset.py:188-198) against the oracle (pinned to
    torchvision and to the reference's own function in tests/test_oracle_preprocess.py): two clips with different
    draws in one call, several op orders, with and without the horizontal flip.  fp32 pipeline -> bf16: values agree
    except where a last-ulp fp32 difference flips the bf16 rounding (<= 0.2 % of the values, by one bf16 step)."""
    cases = [((0, 1, 2, 3), 1.21, 0.83, 1.13, 0.031), ((3, 1, 0, 2), 0.74, 1.27, 0.86, -0.044),
             ((2, 3, 1, 0), 1.05, 0.95, 1.19, 0.05), ((1, 0, 3, 2), 0.9, 1.1, 0.8, -0.05)]
    frames = R.seeded_frames(2 * len(cases), H, W, 51)
    frames[0, :40] = frames[0, :40, :, :1]  # a grey band (max == min in the hue op)
    boxes = torch.tensor([box] * frames.shape[0], dtype=torch.int32, device="cuda")
    rows = _jitter_rows(cases, 2)
    for flip in (False, True):
        got = eng.preprocess_u8(torch.from_numpy(frames).cuda(), boxes, flip_w=flip, jitter=rows)
        got_f = got.float().cpu().numpy()
        for k, c in enumerate(cases):
            want = P.crop_resize_jitter_normalize(frames[2 * k:2 * k + 2], box, *c, flip=flip)
            want_bits = P.to_nhwc4p_bf16_bits(want)
            want_f = (want_bits.astype(np.uint32) << 16).view(np.float32)
            g = got_f[2 * k:2 * k + 2]
            diff = np.abs(g - want_f)
            assert (diff > 0).mean() <= 2e-3, (k, flip, (diff > 0).mean())
            assert diff.max() <= 0.0157, (k, flip, diff.max())  # one bf16 step at |x| < 4
            assert np.array_equal(g[:, :, :4], want_f[:, :, :4]) and np.array_equal(g[..., 3], want_f[..., 3])


def test_extract_with_color_jitter_vs_c_oracle(eng, backbone):
    """Seam B with the jitter variant end to end: features within the trunk's tolerance of the fp32 C oracle fed with
    the oracle's jittered crops."""
    frames = R.seeded_frames(2, 260, 250, 52)
    box = (5, 7, 240, 240)
    case = ((1, 3, 0, 2), 1.18, 0.77, 1.2, -0.02)
    feats = eng.extract_u8(torch.from_numpy(frames).cuda(), torch.tensor([box] * 2, dtype=torch.int32, device="cuda"),
                           jitter=_jitter_rows([case], 2))
    assert eng.launches == 42  # two K1 launches (grey-level row sums, then the fused jitter + normalise) + the trunk
    ref = R.features(P.crop_resize_jitter_normalize(frames, box, *case), R.param_list(backbone))
    err, cos = frame_errors(feats.cpu().numpy(), ref)
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err, cos)
    plain = eng.extract_u8(torch.from_numpy(frames).cuda(), torch.tensor([box] * 2, dtype=torch.int32, device="cuda"))
    assert not torch.equal(plain, feats)
    with pytest.raises(RuntimeError, match="jitter"):
        eng.extract_u8(torch.from_numpy(frames).cuda(), None, jitter=torch.zeros(2, 10, device="cuda"))


def test_preprocess_whole_frame_when_no_boxes(eng):
    frames = R.seeded_frames(2, 224, 224, 23)
    got = eng.preprocess_u8(torch.from_numpy(frames).cuda(), None)
    want = P.to_nhwc4p_bf16_bits(P.crop_resize_normalize(frames, (0, 0, 224, 224)))
    assert np.array_equal(got.view(torch.int16).cpu().numpy().view(np.uint16), want)


# ------------------------------------------------------------------------------------------------ per-layer
def _layer_modules(bb):
    out = {"conv1": (bb[0], bb[1]), "conv1+maxpool": (bb[0], bb[1])}
    for li in range(4):
        for bi, blk in enumerate(bb[4 + li]):
            p = f"layer{li + 1}.{bi}"
            out[p + ".conv1"] = (blk.conv1, blk.bn1)
            out[p + ".conv2"] = (blk.conv2, blk.bn2)
            out[p + ".conv3"] = (blk.conv3, blk.bn3)
            if blk.downsample is not None:
                out[p + ".downsample"] = (blk.downsample[0], blk.downsample[1])
    return out


def _unique_layers(plan):
    seen = set()
    for i, (name, L) in enumerate(zip(plan.names, plan.layers)):
        key = (L.kind, L.cin, L.cout, L.r, L.stride, L.hin, L.res_buf >= 0, L.relu, L.gap, L.in2_buf >= 0)
        if key not in seen:
            seen.add(key)
            yield i, name, L


@pytest.mark.parametrize("n", [1, 5])
def test_every_layer_shape_vs_fp32_torch(eng, backbone, n):
    """Each of the 24 unique GEMM shapes (+ stem, maxpool, gap), on the same bf16-rounded operands, vs fp32 PyTorch.
    n = 1 and 5 exercise the partial last M tile and the odd-frame tail of the fused average pool."""
    mods = _layer_modules(backbone)
    g = torch.Generator(device="cuda").manual_seed(100 + n)
    checked = 0
    for i, name, L in _unique_layers(eng.plan):
        if L.kind == 2:
            x = torch.relu(torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g)).to(torch.bfloat16)
            got = eng.run_layer(i, x)
            ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
            assert torch.equal(got.float(), ref), name
            checked += 1
            continue
        conv, bn = mods[name.replace("+downsample", "")]
        w, b = phdfx.fold_conv_bn(conv, bn)
        w = w.to(torch.bfloat16).float().cuda()
        if L.kind in (1, 3):
            x_nchw = torch.randn(n, 3, 224, 224, device="cuda", generator=g)
            x_in = torch.zeros(n, 224, 232, 4, device="cuda", dtype=torch.bfloat16)
            x_in[:, :, 4:228, :3] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
            x_ref = x_nchw.to(torch.bfloat16).float()
        else:
            x_in = torch.randn(n, L.hin, L.win, L.cin, device="cuda", generator=g).to(torch.bfloat16)
            x_ref = x_in.float().permute(0, 3, 1, 2)
        ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
        res = None
        if L.res_buf >= 0:
            res = torch.randn(n, ho, ho, L.cout, device="cuda", generator=g).to(torch.bfloat16)
        x2 = None
        if L.in2_buf >= 0:
            x2 = torch.randn(n, L.hin2, L.hin2, L.cin2, device="cuda", generator=g).to(torch.bfloat16)
        got = eng.run_layer(i, x_in, res, x2)
        ref = F.conv2d(x_ref, w, b.cuda(), stride=conv.stride, padding=conv.padding)
        if x2 is not None:  # fused down-sample branch: + bn_d(conv_d(x)) (resnet.py:157-160)
            dconv, dbn = mods[name.replace("conv3+downsample", "downsample")]
            wd, bd = phdfx.fold_conv_bn(dconv, dbn)
            ref = ref + F.conv2d(x2.float().permute(0, 3, 1, 2), wd.to(torch.bfloat16).float().cuda(), bd.cuda(),
                                 stride=dconv.stride)
        if res is not None:
            ref = ref + res.float().permute(0, 3, 1, 2)
        if L.relu:
            ref = torch.relu(ref)
        if L.kind == 3:  # fused stem: the conv output is rounded to bf16 before the max-pool
            ref = F.max_pool2d(ref, 3, 2, 1)
        if L.gap:
            ref = ref.mean(dim=(2, 3))
            tol = 1e-4  # fp32 in, fp32 out
        else:
            ref = ref.permute(0, 2, 3, 1)
            tol = 6e-3  # one bf16 rounding of the output
        err = (got.float() - ref).abs().max().item() / ref.abs().max().item()
        assert err < tol, f"{name}: normalised error {err}"
        checked += 1
    assert checked >= 24


def _chain_reference(eng, mods, first, t1, x_or_res):
    """fp32 PyTorch restatement of one fused chain on the same bf16-rounded operands, rounding to bf16 exactly where the
    kernel does (t2, the block output)."""
    names, layers = eng.plan.names, eng.plan.layers
    span = eng.chain_span(first)

    def wb(name):
        conv, bn = mods[name]
        w, b = phdfx.fold_conv_bn(conv, bn)
        return w.to(torch.bfloat16).float().cuda(), b.cuda()

    nchw = lambda t: t.float().permute(0, 3, 1, 2)
    w2, b2 = wb(names[first])
    t2 = torch.relu(F.conv2d(nchw(t1), w2, b2, padding=1)).to(torch.bfloat16).float()
    n3 = names[first + 1]
    w3, b3 = wb(n3.replace("+downsample", ""))
    y = F.conv2d(t2, w3, b3)
    if layers[first + 1].in2_buf >= 0:
        wd, bd = wb(n3.replace("conv3+downsample", "downsample"))
        y = y + F.conv2d(nchw(x_or_res), wd, bd)
    else:
        y = y + nchw(x_or_res)
    out = torch.relu(y)
    t1n = None
    if span == 3:
        w1, b1 = wb(names[first + 2])
        t1n = torch.relu(F.conv2d(out.to(torch.bfloat16).float(), w1, b1)).permute(0, 2, 3, 1)
    return out.permute(0, 2, 3, 1), t1n


@pytest.mark.parametrize("n", [1, 3, 11, 29])
def test_chain_kernel_vs_fp32_torch(eng, backbone, n):
    """bottleneck_chain_kernel (conv2 -> conv3 [+identity | +down-sample] [-> next conv1] in one launch), every
    variant the plan uses (three in layer1, the streamed-conv3 one in layer2), on explicit tensors.  n = 11 / 29 give
    each CTA of a 148-SM grid several tiles (the software pipeline across tiles), n = 1 fewer tiles than SMs."""
    mods = _layer_modules(backbone)
    g = torch.Generator(device="cuda").manual_seed(300 + n)
    firsts = [i for i in range(len(eng.plan.layers)) if eng.chain_span(i) > 0]
    assert [eng.chain_span(i) for i in firsts] == [3, 3, 3, 2, 2, 2]
    for first in firsts:
        L3 = eng.plan.layers[first + 1]
        hw = L3.hin
        t1 = torch.relu(torch.randn(n, hw, hw, L3.cin, device="cuda", generator=g)).to(torch.bfloat16)
        c = L3.cin2 if L3.in2_buf >= 0 else L3.cout
        xr = torch.randn(n, hw, hw, c, device="cuda", generator=g).to(torch.bfloat16)
        out, t1n = eng.run_chain(first, t1, xr)
        ref_out, ref_t1n = _chain_reference(eng, mods, first, t1, xr)
        err = (out.float() - ref_out).abs().max().item() / ref_out.abs().max().item()
        assert err < 6e-3, f"{eng.plan.names[first]}: block output normalised error {err}"
        if ref_t1n is not None:
            err = (t1n.float() - ref_t1n).abs().max().item() / ref_t1n.abs().max().item()
            assert err < 1e-2, f"{eng.plan.names[first + 2]}: next t1 normalised error {err}"


def test_chains_are_bit_identical_to_per_conv_kernels(backbone):
    """The fused chain performs the same MMAs in the same order with the same roundings as the three kernels it
    replaces: features are bit-identical with PHDFX_NO_CHAIN=1, at a batch that gives every CTA several tiles."""
    frames = torch.from_numpy(R.seeded_frames(64, 224, 224, 41)).cuda()
    fused = phdfx.B200Backbone(backbone, device=0, max_frames=64, waves=((0, 0),))
    os.environ["PHDFX_NO_CHAIN"] = "1"
    try:
        plain = phdfx.B200Backbone(backbone, device=0, max_frames=64, waves=((0, 0),))
    finally:
        del os.environ["PHDFX_NO_CHAIN"]
    a = fused.extract_u8(frames, None)
    b = plain.extract_u8(frames, None)
    assert plain.launches - fused.launches == 9  # 6 chains replace 15 launches (both group layer3's CTA-pair convs alike)
    assert torch.equal(a, b)
    fused.close()
    plain.close()


@pytest.mark.parametrize("n", [1, 3])
def test_unfused_stem_and_maxpool_kernels(backbone, n):
    """The separate implicit-GEMM stem and max-pool kernels (fuse_stem_pool=False) and the fused kernel agree on the
    whole path up to fp32 summation order (the fused kernel accumulates the filter rows that feed only one of a step's
    two conv rows first, then the shared ones as N = 128 MMAs)."""
    fused = phdfx.B200Backbone(backbone, device=0, max_frames=4)
    unfused = phdfx.B200Backbone(backbone, device=0, max_frames=4, fuse_stem_pool=False)
    frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 31)).cuda()
    a = fused.extract_u8(frames, None)
    b = unfused.extract_u8(frames, None)
    assert fused.launches == 40 and unfused.launches == 42  # K1 rides in the fused stem kernel only
    err, cos = frame_errors(b.cpu().numpy(), a.cpu().numpy())
    assert err.max() < 5e-3 and cos.min() > 0.99999
    # separate down-sample launches + residual add (rounds the branch to bf16 first): same features within bf16 noise
    plain = phdfx.B200Backbone(backbone, device=0, max_frames=4, fuse_downsample=False)
    c = plain.extract_u8(frames, None)
    assert plain.launches == 46
    err, cos = frame_errors(c.cpu().numpy(), a.cpu().numpy())
    assert err.max() < 1e-2 and cos.min() > 0.9999
    plain.close()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.relu(torch.randn(n, 112, 112, 64, device="cuda", generator=g)).to(torch.bfloat16)
    got = unfused.run_layer(1, x)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(got.float(), ref)
    fused.close()
    unfused.close()


# ------------------------------------------------------------------------------------------------ whole path
def test_config1_features_vs_reference_golden(eng, golden):
    """BASELINE config 1 stand-in: 16 frames (2 clips x 8), 1002x1000 uint8, box (100,200,517,517), Seam B
    (uint8 in, features out) against features the reference's own CPU fp32 path produced."""
    n, H, W, top, left, hh, ww, seed = (int(v) for v in golden["meta"][:8])
    frames = torch.from_numpy(R.seeded_frames(n, H, W, seed)).cuda()
    boxes = torch.tensor([(top, left, hh, ww)] * n, dtype=torch.int32, device="cuda")
    feats = eng.extract_u8(frames, boxes)
    # the reference's reshape at :296
    clips = feats.view(2, 8, -1)
    assert clips.shape == (2, 8, 2048)
    err, cos = frame_errors(feats.cpu().numpy(), golden["feats"])
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err, cos)


def test_identity_size_input_vs_golden(eng, golden):
    frames = torch.from_numpy(R.seeded_frames(4, 224, 224, 2)).cuda()
    feats = eng.extract_u8(frames, None).cpu().numpy()
    err, cos = frame_errors(feats, golden["feats_identity"])
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err, cos)


def test_seam_a_module_call_vs_c_oracle(eng, backbone):
    """Seam A: backbone(x).flatten(1).view(Bv, T, -1) verbatim (src/preprocess_resnet_features.py:295-296), checked
    against the plain-C oracle on 2 frames."""
    frames = R.seeded_frames(2, 300, 280, 5)
    x = P.crop_resize_normalize(frames, (20, 30, 217, 217))
    v_video = torch.from_numpy(x).view(1, 2, 3, 224, 224).cuda()
    Bv, T, C, H, W = v_video.shape
    xx = v_video.view(Bv * T, C, H, W).contiguous()
    out = eng(xx)
    assert out.shape == (2, 2048, 1, 1) and out.dtype == torch.float32
    feats = out.flatten(1).view(Bv, T, -1)
    ref = R.features(x, R.param_list(backbone))
    err, cos = frame_errors(feats.view(2, -1).cpu().numpy(), ref)
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err, cos)


def test_seam_a_equals_seam_b(eng):
    """Normalised-fp32 entry and uint8 entry agree bit-for-bit (both feed the same bf16 NHWC4p bytes)."""
    frames = R.seeded_frames(3, 240, 250, 6)
    box = (4, 6, 230, 230)
    a = eng(torch.from_numpy(P.crop_resize_normalize(frames, box)).cuda()).flatten(1)
    b = eng.extract_u8(torch.from_numpy(frames).cuda(), torch.tensor([box] * 3, dtype=torch.int32, device="cuda"))
    assert torch.equal(a, b)


def test_external_input_buffer_equals_arena(eng):
    frames = torch.from_numpy(R.seeded_frames(3, 224, 224, 7)).cuda()
    x4 = eng.preprocess_u8(frames, None)
    assert torch.equal(eng.forward_nhwc4p(x4), eng.extract_u8(frames, None))


# ------------------------------------------------------------------------------------------------ edge cases / properties
@pytest.mark.parametrize("n", [1, 2, 3, 7, 32])
def test_batch_sizes_and_frame_independence(eng, n):
    """Frames are independent (BN in eval mode): features of a frame do not depend on batch size or position —
    bit-exact, since every output row is produced by the same MMA / reduction order wherever it sits in a tile."""
    frames = torch.from_numpy(R.seeded_frames(32, 224, 224, 8)).cuda()
    full = eng.extract_u8(frames, None)
    sub = eng.extract_u8(frames[:n].contiguous(), None)
    assert torch.equal(sub, full[:n])
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(n)).cuda()
    permuted = eng.extract_u8(frames[perm].contiguous(), None)
    assert torch.equal(permuted, full[perm])


def test_chunking_over_max_frames(eng):
    frames = torch.from_numpy(R.seeded_frames(70, 224, 224, 9)).cuda()  # max_frames = 32 -> 32 + 32 + 6
    feats = eng.extract_u8(frames, None)
    again = torch.cat([eng.extract_u8(frames[i:i + 10].contiguous(), None) for i in range(0, 70, 10)])
    assert torch.equal(feats, again)
    assert not torch.isnan(feats).any() and feats.min() >= 0  # post-ReLU mean


def test_empty_batch(eng):
    """Zero frames in, zero rows out (no launch), on both seams."""
    out = eng.extract_u8(torch.empty(0, 224, 224, 3, dtype=torch.uint8, device="cuda"), None)
    assert tuple(out.shape) == (0, 2048) and eng.launches == 0
    out = eng(torch.empty(0, 3, 224, 224, device="cuda"))
    assert tuple(out.shape) == (0, 2048, 1, 1)


def test_forward_timed_hook(eng):
    """phdfx_forward_timed (in-situ per-launch timing, BASELINE config 3): same features as the plain call, one
    positive time per launch, fused chains reported as one entry."""
    frames = torch.from_numpy(R.seeded_frames(6, 224, 224, 15)).cuda()
    x4 = eng.preprocess_u8(frames, None)
    feats, times = eng.forward_timed(x4)
    assert torch.equal(feats, eng.forward_nhwc4p(x4))
    assert len(times) == 40 and all(ms > 0 for _, ms in times)
    names = [nm for nm, _ in times]
    assert names[0] == "conv1+maxpool" and "layer1.1.conv2+layer1.1.conv3+layer1.2.conv1" in names
    assert "layer2.1.conv2+layer2.1.conv3" in names and names[-1] == "layer4.2.conv3"


def test_deterministic(eng):
    frames = torch.from_numpy(R.seeded_frames(9, 224, 224, 10)).cuda()
    a = eng.extract_u8(frames, None).clone()
    b = eng.extract_u8(frames, None)
    assert torch.equal(a, b)


def test_time_reverse_variant_is_a_permutation(eng):
    """The reference's `trev` augmentation (src/dataset.py:201-207) reverses the clip in time; frames being
    independent, its features are exactly the reversed `orig` features."""
    frames = torch.from_numpy(R.seeded_frames(8, 224, 224, 11)).cuda()
    orig = eng.extract_u8(frames, None)
    trev = eng.extract_u8(torch.flip(frames, dims=[0]).contiguous(), None)
    assert torch.equal(trev, torch.flip(orig, dims=[0]))


def test_streaming_host_api(eng):
    """Host-buffer front door: pinned H2D -> K1 -> trunk -> D2H, double-buffered; equals the device-resident path."""
    frames = R.seeded_frames(50, 224, 224, 12)
    box = (0, 0, 224, 224)
    se = phdfx.StreamingExtractor(eng, batch=16)
    host = se(torch.from_numpy(frames), torch.tensor([box] * 50))
    dev = eng.extract_u8(torch.from_numpy(frames).cuda(), None).cpu()
    assert host.shape == (50, 2048) and torch.equal(host, dev)
    assert se.h2d_bytes == 50 * 224 * 224 * 3 + 50 * 16 and se.d2h_bytes == 50 * 2048 * 4
    se16 = phdfx.StreamingExtractor(eng, batch=16, save_fp16=True)  # --save-fp16 (:146)
    assert torch.equal(se16(torch.from_numpy(frames)), dev.to(torch.float16))


def test_cuda_graph_replay_is_bit_identical(eng):
    """One whole step captured into a CUDA graph (programmatic-dependent-launch edges included): refill the input
    buffers in place, replay, same bits as the eager call."""
    frames = torch.from_numpy(R.seeded_frames(20, 240, 250, 14)).cuda()
    boxes = torch.tensor([(3, 4, 230, 230)] * 20, dtype=torch.int32, device="cuda")
    buf = frames[:10].clone()
    g = eng.capture_extract(buf, boxes[:10].clone())
    assert torch.equal(g.replay(), eng.extract_u8(frames[:10].contiguous(), boxes[:10].contiguous()))
    buf.copy_(frames[10:])
    assert torch.equal(g.replay(), eng.extract_u8(frames[10:].contiguous(), boxes[10:].contiguous()))
    assert g.launches == 41  # 240x250 frames: K1 stays a launch of its own (api.cu: stem_can_fuse_k1)


def test_errors_are_loud(eng):
    with pytest.raises(RuntimeError):
        eng(torch.zeros(2, 3, 200, 200, device="cuda"))
    with pytest.raises(RuntimeError):
        eng.extract_u8(torch.zeros(2, 224, 224, 3, dtype=torch.uint8), None)  # host tensor
    with pytest.raises(RuntimeError):
        eng.forward_nhwc4p(torch.zeros(33, 224, 232, 4, device="cuda", dtype=torch.bfloat16))  # > max_frames
    with pytest.raises(RuntimeError, match="residual"):
        i = eng.plan.names.index("layer1.1.conv3")
        eng.run_layer(i, torch.zeros(1, 56, 56, 64, device="cuda", dtype=torch.bfloat16), None)
    with pytest.raises(RuntimeError, match="second input"):
        j = eng.plan.names.index("layer1.0.conv3+downsample")
        eng.run_layer(j, torch.zeros(1, 56, 56, 64, device="cuda", dtype=torch.bfloat16), None)


def test_full_batch_256_properties():
    """BASELINE config 2 size (batch 256): too big for the CPU oracle, so check size-independent properties —
    equality with the same frames pushed through in small batches, and the native library is what ran."""
    bb = R.seeded_backbone()
    e = phdfx.B200Backbone(bb, device=0, max_frames=256)
    raw = R.seeded_frames(256, 224, 224, 13)
    frames = torch.from_numpy(raw).cuda()
    big = e.extract_u8(frames, None)  # the default frame-wave schedule
    waved_launches = e.launches
    e.set_waves(((0, 0),))
    whole = e.extract_u8(frames, None)
    # un-waved: fused K1/stem/maxpool + 28 conv launches (4 down-samples ride in conv3; layer1's conv2 -> conv3 ->
    # next conv1 chains and layer2's conv2 -> conv3 chains are one launch each; in layer3 / layer4 every block's conv1 ->
    # conv2 [-> conv3 + down-sample] is one multi-phase CTA-pair launch), all ours
    assert e.launches == 29 and waved_launches >= 29
    assert torch.equal(big, whole)
    small = torch.cat([e.extract_u8(frames[i:i + 37].contiguous(), None) for i in range(0, 256, 37)])
    assert torch.equal(big, small)
    assert torch.isfinite(big).all()
    # and directly against the oracle for a handful of the 256 frames (first, last, and two from the middle waves)
    pick = [0, 100, 201, 255]
    x = np.stack([P.crop_resize_normalize(raw[i:i + 1], (0, 0, 224, 224))[0] for i in pick])
    ref = R.features(x, R.param_list(bb))
    err, cos = frame_errors(big[pick].cpu().numpy(), ref)
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err, cos)
    e.close()


def test_config5_fixture_is_what_head_extracts(golden_dir):
    """BASELINE config 5 is checked on CPU (tests/test_e2e_loss.py) against features a B200 extracted earlier
    (tests/golden/e2e_b200_feats.npz, tools/e2e_extract.py).  Re-extract the same clips with the library under test:
    within the north-star tolerance always, and bit for bit when the fixture was made from these very kernel sources."""
    from phdfx.synthetic import SyntheticH36MClips, csrc_sha

    fx = np.load(os.path.join(golden_dir, "e2e_b200_feats.npz"))
    n_clips, seq_len, H, W, side, seed = (int(v) for v in fx["cfg"])
    ds = SyntheticH36MClips(n_clips, seq_len=seq_len, height=H, width=W, subjects=(1, 6, 7, 8), seed=seed,
                            box_side=side)
    e = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=n_clips * seq_len)
    fr = torch.stack([ds.frames(i) for i in range(n_clips)]).view(n_clips * seq_len, H, W, 3).cuda()
    bx = torch.stack([ds.box(i) for i in range(n_clips)]).to(torch.int32).repeat_interleave(seq_len, dim=0).cuda()
    got = e.extract_u8(fr, bx).cpu().numpy()
    want = fx["feats"].reshape(-1, 2048)
    err, cos = frame_errors(got, want)
    assert err.max() <= NORM_TOL and cos.min() >= COS_TOL, (err.max(), cos.min())
    same_sources = "csrc_sha" in fx and str(fx["csrc_sha"]) == csrc_sha()
    if same_sources:
        assert np.array_equal(got, want), "fixture was extracted from these sources but the bits differ"
    print(f"config-5 fixture vs HEAD: norm_err {err.max():.2e} min cos {cos.min():.7f} "
          f"bit_identical {np.array_equal(got, want)} same_sources {same_sources}")
    e.close()


@pytest.mark.parametrize("switch,exact", [("PHDFX_NO_CG2", True), ("PHDFX_NO_REV", True), ("PHDFX_NO_SMALL_N", True),
                                          ("PHDFX_NO_FUSE_K1", True), ("PHDFX_NO_HALO", False)])
def test_kernel_selection_switches(backbone, switch, exact):
    """The A/B switches select other kernels for the same layers (1-CTA instead of CTA-pair implicit GEMM, ascending tile
    order, 256-wide N tiles for small launches, im2col instead of the halo patch mode).  They are read per handle at
    phdfx_create.  All but NO_HALO keep every output element's K order, hence bit-identical features; the im2col path
    walks K tap-major instead of channel-block-major, so layer2's 3x3 convs round differently (fp32 accumulation)."""
    base = phdfx.B200Backbone(backbone, device=0, max_frames=40)
    os.environ[switch] = "1"
    try:
        alt = phdfx.B200Backbone(backbone, device=0, max_frames=40)
    finally:
        del os.environ[switch]
    again = phdfx.B200Backbone(backbone, device=0, max_frames=40)  # created after the switch was removed: default kernels
    for n in (1, 3, 8, 37):
        frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 50 + n)).cuda()
        a, b, c = base.extract_u8(frames, None), alt.extract_u8(frames, None), again.extract_u8(frames, None)
        assert torch.equal(a, c)
        if exact:
            assert torch.equal(a, b), (switch, n)
        else:
            err, cos = frame_errors(b.cpu().numpy(), a.cpu().numpy())
            assert err.max() < 5e-3 and cos.min() > 0.99999, (switch, n, err.max())
    for e in (base, alt, again):
        e.close()


@pytest.mark.parametrize("H,W,box,flip,n", [
    (224, 224, None, False, 9),                 # identity-size frames: the table-lookup path
    (224, 224, None, True, 5),                  # mirrored read
    (300, 280, (20, 30, 217, 217), False, 7),   # bilinear, odd byte alignment of the crop rows
    (241, 263, "ragged", True, 13),             # a different box per frame, mirrored
    (1002, 1000, (100, 200, 517, 517), False, 3),   # H36M-sized frames (3 KB staging rows)
    (64, 48, None, False, 6),                   # 4.7x upscale: many rows share source rows
])
def test_k1_inside_the_stem_equals_k1_as_a_launch(backbone, H, W, box, flip, n):
    """extract_u8 runs K1 inside the stem kernel's converter warps (stem_pool_sm100.cuh, FUSE_K1) with the very device
    function preprocess_u8_kernel uses (pinned bit for bit against the reference's own crop/resize/normalise above): the
    features equal those of K1 as a launch of its own (PHDFX_NO_FUSE_K1=1) bit for bit, for every crop geometry, also
    in frame waves (the fused launch then starts at a frame offset of the caller's buffer) and at the first / last
    vector of the buffer (frames sliced out of a larger allocation at odd byte offsets)."""
    os.environ["PHDFX_FUSE_K1"] = "1"  # for every geometry (the default only fuses where it is not slower, api.cu)
    try:
        fused = phdfx.B200Backbone(backbone, device=0, max_frames=16)
    finally:
        del os.environ["PHDFX_FUSE_K1"]
    os.environ["PHDFX_NO_FUSE_K1"] = "1"
    try:
        plain = phdfx.B200Backbone(backbone, device=0, max_frames=16)
    finally:
        del os.environ["PHDFX_NO_FUSE_K1"]
    raw = torch.from_numpy(R.seeded_frames(n, H, W, 900 + H + n)).cuda()
    boxes = None
    if box == "ragged":
        g = np.random.default_rng(H + W)
        rows = []
        for _ in range(n):
            side = int(g.integers(40, min(H, W)))
            rows.append((int(g.integers(0, H - side + 1)), int(g.integers(0, W - side + 1)), side, side))
        boxes = torch.tensor(rows, dtype=torch.int32, device="cuda")
    elif box is not None:
        boxes = torch.tensor([box] * n, dtype=torch.int32, device="cuda")
    a = fused.extract_u8(raw, boxes, flip_w=flip)
    assert fused.launches == 40
    b = plain.extract_u8(raw, boxes, flip_w=flip)
    assert plain.launches == 41
    assert torch.equal(a, b)
    # frames that do not start at an aligned address / end at the end of an allocation
    pad = torch.zeros(raw.numel() + 7, dtype=torch.uint8, device="cuda")
    pad[7:] = raw.flatten()
    shifted = pad[7:].view(n, H, W, 3)
    assert torch.equal(fused.extract_u8(shifted, boxes, flip_w=flip), a)
    fused.set_waves(((0, 4), (7, 0)))
    assert torch.equal(fused.extract_u8(raw, boxes, flip_w=flip), a)
    fused.close()
    plain.close()


@pytest.mark.parametrize("n", [64, 100, 200, 256])
def test_multi_phase_launches_are_bit_identical(backbone, n):
    """conv1 -> conv2 [-> conv3 + down-sample] of a layer3 / layer4 block run as ONE multi-phase CTA-pair launch
    (csrc/conv_igemm_cg2_multi_sm100.cuh): every tile is computed as by the per-conv kernel, the phases meet on per-frame
    progress counters.  Features equal those of a handle created with PHDFX_NO_MULTI=1 bit for bit — repeated plain
    launches (a missed dependency would be a race) and graph replay."""
    multi = phdfx.B200Backbone(backbone, device=0, max_frames=n)
    os.environ["PHDFX_NO_MULTI"] = "1"
    try:
        plain = phdfx.B200Backbone(backbone, device=0, max_frames=n)
    finally:
        del os.environ["PHDFX_NO_MULTI"]
    frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 500 + n)).cuda()
    want = plain.extract_u8(frames, None).clone()
    assert plain.launches == 40
    out = torch.empty(n, 2048, device="cuda")
    for _ in range(6):
        out.fill_(float("nan"))
        multi.extract_u8(frames, None, out=out)
        assert torch.equal(out, want), n
    assert multi.launches == (29 if n >= 100 else 33), multi.launches  # at n = 64 layer4's convs use narrow 1-CTA tiles
    g = multi.capture_extract(frames, None)
    for _ in range(6):
        assert torch.equal(g.replay(), want), n
    # per-launch timing keeps one launch per conv (and the same bits)
    assert len(multi.forward_timed(multi.preprocess_u8(frames, None))[1]) == 40
    multi.close()
    plain.close()


@pytest.mark.parametrize("n", [100, 200, 256])
def test_frame_progress_links_are_bit_identical(backbone, n):
    """Launches that start on their predecessor's per-frame progress counters instead of waiting for its whole grid
    (include/phdfx.h: phdfx_linked_launches; csrc/conv_igemm_sm100.cuh) read the same bytes in the same order:
    features of a handle created with PHDFX_FLAGS=1 equal those of a default handle, under plain launches (repeated — a
    missed dependency would be a race) and under graph replay.  Links need full grids: n = 100 links layer3 only,
    n >= 194 layer4 as well."""
    plain = phdfx.B200Backbone(backbone, device=0, max_frames=n)
    os.environ["PHDFX_FLAGS"] = "1"
    try:
        base = phdfx.B200Backbone(backbone, device=0, max_frames=n)
    finally:
        del os.environ["PHDFX_FLAGS"]
    assert plain.linked_launches(n) == 0
    links = base.linked_launches(n)
    if links == 0:
        base.close()
        plain.close()
        pytest.skip("libphdfx.so was built without PHDFX_EXPERIMENTAL=1: no counter code in the single-launch kernels")
    assert links >= (16 if n >= 194 else 10), links
    assert base.linked_launches(8) == 0  # small grids: more than two launches could be resident at once
    frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 300 + n)).cuda()
    want = plain.extract_u8(frames, None).clone()
    out = torch.empty(n, 2048, device="cuda")
    for _ in range(6):
        out.fill_(float("nan"))
        base.extract_u8(frames, None, out=out)
        assert torch.equal(out, want), n
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        base.extract_u8(frames, None, out=out)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        base.extract_u8(frames, None, out=out)
    for _ in range(6):
        out.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want), n
    del g
    base.close()
    plain.close()


def test_small_batches_are_bit_identical_to_their_rows_in_a_big_batch(backbone):
    """Launches with few tiles use narrower N tiles (api.cu: geometry): a frame's features must not depend on it."""
    e = phdfx.B200Backbone(backbone, device=0, max_frames=64)
    frames = torch.from_numpy(R.seeded_frames(64, 224, 224, 77)).cuda()
    big = e.extract_u8(frames, None)
    for n, at in ((1, 0), (1, 63), (2, 10), (5, 31), (16, 40)):
        assert torch.equal(e.extract_u8(frames[at:at + n].contiguous(), None), big[at:at + n]), (n, at)
    e.close()
