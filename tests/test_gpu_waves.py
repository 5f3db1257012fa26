"""Frame-wave schedules (phdfx_set_schedule): every schedule must give bit-identical features — a frame meets the same
kernels, tiles and K order whatever wave it travels in — and unsafe schedules must be refused, not run."""
import numpy as np
import pytest
import torch

import phdfx
import resnet50_ref as R

pytestmark = pytest.mark.gpu

SCHEDULES = [
    ((0, 32), (7, 0)),
    ((0, 37), (7, 0)),
    ((0, 21), (3, 42), (7, 0)),
    ((0, 16), (3, 0)),
    ((0, 24), (7, 100), (13, 0)),
    ((0, 19), (3, 19), (7, 19), (13, 19)),
]


@pytest.fixture(scope="module")
def eng256():
    e = phdfx.B200Backbone(R.seeded_backbone(), device=0, max_frames=256)
    yield e
    e.close()


@pytest.mark.parametrize("n", [256, 77])
def test_every_schedule_is_bit_identical(eng256, n):
    e = eng256
    frames = torch.from_numpy(R.seeded_frames(n, 240, 260, 21 + n)).cuda()
    rng = np.random.default_rng(n)
    boxes = torch.tensor([[int(rng.integers(0, 20)), int(rng.integers(0, 30)), 200 + int(rng.integers(0, 20)),
                           200 + int(rng.integers(0, 20))] for _ in range(n)], dtype=torch.int32, device="cuda")
    e.set_waves(((0, 0),))
    ref = e.extract_u8(frames, boxes).clone()
    base_launches = e.launches
    x4 = e.preprocess_u8(frames, boxes)
    for sched in SCHEDULES:
        for reuse in (True, False):
            e.set_waves(sched, reuse=reuse)
            got = e.extract_u8(frames, boxes)
            assert torch.equal(got, ref), (sched, reuse, "extract_u8")
            assert e.launches > base_launches
            got = e.forward_nhwc4p(x4)  # caller-owned input tensor: addressed by absolute frame number
            assert torch.equal(got, ref), (sched, reuse, "forward_nhwc4p")
            got = e.extract_u8(frames, boxes, flip_w=True)
            e.set_waves(((0, 0),))
            assert torch.equal(got, e.extract_u8(frames, boxes, flip_w=True)), (sched, reuse, "flip")
    e.set_waves(((0, 0),))


def test_waved_graph_replay_and_timed_forward(eng256):
    e = eng256
    n = 256
    frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 9)).cuda()
    e.set_waves(((0, 0),))
    ref = e.extract_u8(frames, None).clone()
    e.set_waves(((0, 32), (7, 0)))
    g = e.capture_extract(frames, None)
    g.out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(g.out, ref)
    feats, times = e.forward_timed(e.preprocess_u8(frames, None))
    assert torch.equal(feats, ref)
    assert len(times) == 40 and all(ms > 0 for _, ms in times)  # one figure per launch of the un-waved list
    e.set_waves(((0, 0),))


def test_jitter_variant_under_waves(eng256):
    e = eng256
    n = 80
    frames = torch.from_numpy(R.seeded_frames(n, 224, 224, 17)).cuda()
    rows = torch.stack([phdfx.jitter_params((1, 3, 0, 2), 1.0 + 0.002 * i, 0.8 + 0.001 * i, 1.1, 0.01)
                        for i in range(n)]).cuda()
    e.set_waves(((0, 0),))
    ref = e.extract_u8(frames, None, jitter=rows).clone()
    e.set_waves(((0, 32), (7, 0)))
    assert torch.equal(e.extract_u8(frames, None, jitter=rows), ref)
    e.set_waves(((0, 0),))


def test_unsafe_schedules_are_refused(eng256):
    e = eng256
    with pytest.raises(RuntimeError, match="inside a fused"):
        i = e.plan.names.index("layer1.1.conv1")  # rides in the previous block's chain launch
        e.set_schedule([0, i], [32, 0])
    with pytest.raises(RuntimeError, match="ascend"):
        e.set_schedule([0, 11, 11], [32, 0, 0])
    with pytest.raises(RuntimeError, match="max_frames"):
        e.set_schedule([0], [100000])
    # a plan without dedicated buffer ids for the tensors crossing the cut: wave-local reuse would clobber them
    bb = R.seeded_backbone()
    import phdfx.weights as Wt
    orig = Wt.build_plan
    try:
        phdfx.backbone.build_plan = lambda *a, **k: orig(*a, **{**k, "stage_after_blocks": ()})
        e2 = phdfx.B200Backbone(bb, device=0, max_frames=64)
    finally:
        phdfx.backbone.build_plan = orig
    frames = torch.from_numpy(R.seeded_frames(64, 224, 224, 3)).cuda()
    ref = e2.extract_u8(frames, None).clone()
    for reuse in (True, False):  # intermediates of one wave would land on live frames of another, in either mode
        with pytest.raises(RuntimeError, match="buffers of their own|scratch"):
            e2.set_waves(((0, 16), (7, 0)), reuse=reuse)
    assert torch.equal(e2.extract_u8(frames, None), ref)  # the refused call left the schedule alone
    assert torch.equal(e2.extract_u8(frames, None), eng256.extract_u8(frames, None))  # cut and un-cut plans agree
    e2.close()
