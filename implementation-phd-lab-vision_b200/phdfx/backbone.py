"""B200Backbone — the object that takes the place of the reference's `backbone` (src/preprocess_resnet_features.py:207-209).

Seam A (module-shaped):  feats = backbone(x).flatten(1).view(Bv, T, -1)   (:296) works verbatim:
    backbone(x: cuda fp32 [N,3,224,224], ImageNet-normalised) -> cuda fp32 [N,2048,1,1]
Seam B (fused, fast):    backbone.extract_u8(frames: cuda uint8 [N,H,W,3], boxes int32 [N,4] | None) -> [N,2048]
    which also replaces the CPU crop/resize/normalise of src/dataset.py:141-152,242-245.

All math runs in libphdfx.so (hand-written sm_100a kernels) on the current CUDA stream.  There is no fallback:
constructing this object without a CUDA sm_100 device, or without the built extension, raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .weights import Plan, build_plan


def jitter_params(fn_idx, brightness: Optional[float] = None, contrast: Optional[float] = None,
                  saturation: Optional[float] = None, hue: Optional[float] = None, *,
                  brightness_factor: Optional[float] = None, contrast_factor: Optional[float] = None,
                  saturation_factor: Optional[float] = None, hue_factor: Optional[float] = None) -> torch.Tensor:
    """One row of K1's colour-jitter parameters from what torchvision's v2 ColorJitter.make_params returns
    (transforms/v2/_color.py:146-154: fn_idx = randperm(4), the four factors): fp32 [12].  Accepts the dict as is:
    jitter_params(**ColorJitter(0.3, 0.3, 0.2, 0.05).make_params([]))."""
    vals = [brightness if brightness is not None else brightness_factor,
            contrast if contrast is not None else contrast_factor,
            saturation if saturation is not None else saturation_factor, hue if hue is not None else hue_factor]
    if any(v is None for v in vals):
        raise ValueError("all four factors are required (ColorJitter with every range set, as the reference's)")
    order = [float(int(v)) for v in fn_idx]
    if sorted(order) != [0.0, 1.0, 2.0, 3.0]:
        raise ValueError(f"fn_idx must be a permutation of 0..3, got {list(fn_idx)}")
    b, c, s, h = (float(v) for v in vals)
    return torch.tensor(order + [b, c, 1.0 - c, s, 1.0 - s, h, 0.0, 0.0], dtype=torch.float64).to(torch.float32)


class B200Backbone:
    FEAT_DIM = _lib.FEAT_DIM

    # default frame-wave schedule (phdfx_set_schedule): (first block of the stage | 0 = from the stem, frames per wave)
    DEFAULT_WAVES = ((0, 0),)

    def __init__(self, backbone: nn.Module, device: "int | str | torch.device" = 0, max_frames: int = 1280,
                 fuse_stem_pool: bool = True, fuse_downsample: bool = True, waves=None, reuse: bool = True):
        self._lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("B200Backbone needs a CUDA device (sm_100); this backend has no CPU fallback")
        dev = torch.device(device if not isinstance(device, int) else f"cuda:{device}")
        if dev.type != "cuda":
            raise RuntimeError(f"B200Backbone cannot run on {dev}; this backend has no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.max_frames = int(max_frames)
        self.plan: Plan = build_plan(backbone, fuse_stem_pool=fuse_stem_pool, fuse_downsample=fuse_downsample)
        self._h = C.c_void_p()
        _lib.check(self._lib.phdfx_create(C.byref(self._h), self.device.index, self.max_frames))
        arr = (_lib.LayerDesc * len(self.plan.layers))(*self.plan.layers)
        w, b = self.plan.weights, self.plan.bias
        _lib.check(
            self._lib.phdfx_load_weights(self._h, w.data_ptr(), w.numel(), b.data_ptr(), b.numel(), arr,
                                         len(self.plan.layers)), self._h)
        self.set_waves(self.DEFAULT_WAVES if waves is None else waves, reuse=reuse)

    # ---- execution schedule (include/phdfx.h: phdfx_set_schedule) -----------------------------------------------
    def stage_start(self, after_block: int) -> int:
        """Execution-list index at which a stage that begins behind bottleneck block `after_block` (1-based; 0 = the
        stem) starts: the next block's conv1, or its conv2 when that conv1 rides along in the previous block's fused
        chain launch."""
        if after_block <= 0:
            return 0
        i = self.plan.block_first[after_block]
        prev_conv2 = self.plan.block_first[after_block - 1] + 1
        if self.chain_span(prev_conv2) == 3:
            i += 1
        return i

    def set_waves(self, waves, reuse: bool = True):
        """waves: sequence of (after_block, frames_per_wave): the stage that starts behind bottleneck block
        `after_block` (0 = at the stem; cuts must be in the plan's stage_after_blocks) runs in waves of that many
        frames (0 = whole call).  E.g. ((0, 32), (7, 0)): stem + layer1 + layer2 in 32-frame waves, the rest at once."""
        waves = [(int(a), int(w)) for a, w in waves]
        first = (C.c_int32 * len(waves))(*[self.stage_start(a) for a, _ in waves])
        wv = (C.c_int32 * len(waves))(*[min(w, self.max_frames) for _, w in waves])
        _lib.check(self._lib.phdfx_set_schedule(self._h, first, wv, len(waves), _lib.SCHED_REUSE if reuse else 0),
                   self._h)
        self.waves, self.reuse = tuple(waves), bool(reuse)

    def set_schedule(self, first_layers, wave_frames, reuse: bool = True):
        """Raw form: execution-list indices and wave sizes, as phdfx_set_schedule takes them."""
        first = (C.c_int32 * len(first_layers))(*[int(v) for v in first_layers])
        wv = (C.c_int32 * len(wave_frames))(*[int(v) for v in wave_frames])
        _lib.check(self._lib.phdfx_set_schedule(self._h, first, wv, len(first_layers),
                                                _lib.SCHED_REUSE if reuse else 0), self._h)

    def get_schedule(self):
        first, wv, fl = (C.c_int32 * 16)(), (C.c_int32 * 16)(), C.c_int32()
        n = self._lib.phdfx_get_schedule(self._h, first, wv, 16, C.byref(fl))
        return [(int(first[i]), int(wv[i])) for i in range(n)], int(fl.value)

    # ---- nn.Module-shaped surface the reference script touches (:209) ---------------------------------------
    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.phdfx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check_dev(self, t: torch.Tensor, name: str):
        if not t.is_cuda or t.device.index != self.device.index:
            raise RuntimeError(f"{name} must live on {self.device}, got {t.device}")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")

    def linked_launches(self, n: int) -> int:
        """Launches of a pass over n frames that follow their predecessor's frame progress counters (include/phdfx.h)."""
        return int(self._lib.phdfx_linked_launches(self._h, int(n)))

    @property
    def last_launch_count(self) -> int:
        return int(self._lib.phdfx_last_launch_count(self._h))

    # ---- Seam A ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """x: fp32 (or bf16/fp16, cast) NCHW [N,3,224,224] on the device -> fp32 [N,2048,1,1]."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (3, _lib.IMG, _lib.IMG):
            raise RuntimeError(f"expected [N,3,224,224], got {tuple(x.shape)}")
        x = x.to(torch.float32).contiguous()
        self._check_dev(x, "x")
        n = x.shape[0]
        feats = torch.empty(n, self.FEAT_DIM, device=self.device, dtype=torch.float32)
        launches = 0
        with torch.cuda.device(self.device):
            for i in range(0, n, self.max_frames):
                m = min(self.max_frames, n - i)
                st = self._stream()
                _lib.check(self._lib.phdfx_nchw_f32_to_nhwc_bf16(self._h, x[i:i + m].data_ptr(), m, None, st),
                           self._h)
                launches += self.last_launch_count
                _lib.check(self._lib.phdfx_forward(self._h, None, m, feats[i:i + m].data_ptr(), st), self._h)
                launches += self.last_launch_count
        self.launches = launches
        return feats.view(n, self.FEAT_DIM, 1, 1)

    forward = __call__

    # ---- Seam B ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def extract_u8(self, frames: torch.Tensor, boxes: Optional[torch.Tensor] = None, flip_w: bool = False,
                   out: Optional[torch.Tensor] = None, jitter: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frames: uint8 [N,H,W,3] on the device; boxes: int32 [N,4] (top,left,h,w) or None -> fp32 [N,2048].
        jitter: fp32 [N,12] on the device (see jitter_params / include/phdfx.h): the reference's colour-jitter variant."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise RuntimeError(f"expected uint8 [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}")
        self._check_dev(frames, "frames")
        n, H, W, _ = frames.shape
        if boxes is not None:
            if boxes.dtype != torch.int32 or tuple(boxes.shape) != (n, 4):
                raise RuntimeError("boxes must be int32 [N,4] (top, left, h, w)")
            self._check_dev(boxes, "boxes")
        feats = out if out is not None else torch.empty(n, self.FEAT_DIM, device=self.device, dtype=torch.float32)
        if out is not None:
            if out.dtype != torch.float32 or tuple(out.shape) != (n, self.FEAT_DIM):
                raise RuntimeError("out must be fp32 [N,2048]")
            self._check_dev(out, "out")
        if jitter is not None:
            self._check_jitter(jitter, n)
        launches = 0
        with torch.cuda.device(self.device):
            for i in range(0, n, self.max_frames):
                m = min(self.max_frames, n - i)
                bp = boxes[i:i + m].data_ptr() if boxes is not None else None
                if jitter is None:
                    rc = self._lib.phdfx_extract_u8(self._h, frames[i:i + m].data_ptr(), m, H, W, bp, int(flip_w),
                                                    feats[i:i + m].data_ptr(), self._stream())
                else:
                    rc = self._lib.phdfx_extract_u8_jitter(self._h, frames[i:i + m].data_ptr(), m, H, W, bp,
                                                           int(flip_w), jitter[i:i + m].data_ptr(),
                                                           feats[i:i + m].data_ptr(), self._stream())
                _lib.check(rc, self._h)
                launches += self.last_launch_count
        self.launches = launches
        return feats

    def capture(self, fn) -> "CapturedCall":
        """Capture an arbitrary sequence of engine calls on fixed tensors (fn()) into a CUDA graph."""
        return CapturedCall(self, fn)

    def capture_extract(self, frames: torch.Tensor, boxes: Optional[torch.Tensor] = None, flip_w: bool = False,
                        out: Optional[torch.Tensor] = None) -> "ExtractGraph":
        """Capture extract_u8 on exactly these tensors into a CUDA graph (its 29-42 launches -> one graph launch)."""
        return ExtractGraph(self, frames, boxes, flip_w, out)

    def _check_jitter(self, jitter: torch.Tensor, n: int):
        if jitter.dtype != torch.float32 or tuple(jitter.shape) != (n, _lib.JITTER_FLOATS):
            raise RuntimeError(f"jitter must be fp32 [N,{_lib.JITTER_FLOATS}] (phdfx.jitter_params), got "
                               f"{jitter.dtype} {tuple(jitter.shape)}")
        self._check_dev(jitter, "jitter")

    @torch.no_grad()
    def preprocess_u8(self, frames: torch.Tensor, boxes: Optional[torch.Tensor] = None,
                      flip_w: bool = False, out: Optional[torch.Tensor] = None,
                      jitter: Optional[torch.Tensor] = None) -> torch.Tensor:
        """K1 alone: uint8 [N,H,W,3] -> bf16 NHWC4p [N,224,232,4] (the trunk's input layout); jitter as extract_u8."""
        self._check_dev(frames, "frames")
        n, H, W, _ = frames.shape
        if out is None:
            out = torch.empty(n, _lib.IMG, _lib.IN_WPAD, _lib.IN_CPAD, device=self.device, dtype=torch.bfloat16)
        elif out.dtype != torch.bfloat16 or tuple(out.shape) != (n, _lib.IMG, _lib.IN_WPAD, _lib.IN_CPAD):
            raise RuntimeError("out must be bf16 [N,224,232,4]")
        else:
            self._check_dev(out, "out")
        if jitter is not None:
            self._check_jitter(jitter, n)
        with torch.cuda.device(self.device):
            for i in range(0, n, self.max_frames):
                m = min(self.max_frames, n - i)
                bp = boxes[i:i + m].data_ptr() if boxes is not None else None
                if jitter is None:
                    rc = self._lib.phdfx_preprocess_u8(self._h, frames[i:i + m].data_ptr(), m, H, W, bp, int(flip_w),
                                                       out[i:i + m].data_ptr(), self._stream())
                else:
                    rc = self._lib.phdfx_preprocess_u8_jitter(self._h, frames[i:i + m].data_ptr(), m, H, W, bp,
                                                              int(flip_w), jitter[i:i + m].data_ptr(),
                                                              out[i:i + m].data_ptr(), self._stream())
                _lib.check(rc, self._h)
        return out

    @torch.no_grad()
    def forward_nhwc4p(self, x: torch.Tensor) -> torch.Tensor:
        """Trunk on an explicit NHWC4p bf16 input [N,224,232,4] -> fp32 [N,2048]."""
        self._check_dev(x, "x")
        n = x.shape[0]
        if x.dtype != torch.bfloat16 or tuple(x.shape[1:]) != (_lib.IMG, _lib.IN_WPAD, _lib.IN_CPAD):
            raise RuntimeError("expected bf16 [N,224,232,4]")
        if n > self.max_frames:
            raise RuntimeError(f"n = {n} > max_frames = {self.max_frames}")
        feats = torch.empty(n, self.FEAT_DIM, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.phdfx_forward(self._h, x.data_ptr(), n, feats.data_ptr(), self._stream()), self._h)
        return feats

    @torch.no_grad()
    def forward_timed(self, x: torch.Tensor):
        """Trunk on an NHWC4p bf16 input with every launch timed in situ (CUDA events between launches).
        Returns (features, [(name, ms), ...]); a fused chain is one entry named after the layers it covers."""
        self._check_dev(x, "x")
        n = x.shape[0]
        feats = torch.empty(n, self.FEAT_DIM, device=self.device, dtype=torch.float32)
        cap = len(self.plan.layers)
        ms = (C.c_float * cap)()
        with torch.cuda.device(self.device):
            cnt = self._lib.phdfx_forward_timed(self._h, x.data_ptr(), n, feats.data_ptr(), self._stream(), ms, cap)
        if cnt < 0:
            _lib.check(cnt, self._h)
        names, i = [], 0
        while i < len(self.plan.layers):
            span = max(1, self.chain_span(i))
            names.append("+".join(self.plan.names[i:i + span]))
            i += span
        assert len(names) == cnt, (len(names), cnt)
        return feats, list(zip(names, [float(v) for v in ms[:cnt]]))

    # ---- per-layer hook --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def run_layer(self, layer_id: int, x: torch.Tensor, residual: Optional[torch.Tensor] = None,
                  x2: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run one entry of the execution list on explicit tensors (NHWC bf16; NHWC4p for the stem); x2 = the second
        input of a fused conv3+downsample layer."""
        L = self.plan.layers[layer_id]
        n = x.shape[0]
        self._check_dev(x, "x")
        if L.kind == _lib.PHDFX_STEM_POOL:
            out = torch.empty(n, 56, 56, L.cout, device=self.device, dtype=torch.bfloat16)
        elif L.kind == _lib.PHDFX_MAXPOOL:
            ho = (L.hin + 2 - 3) // 2 + 1
            out = torch.empty(n, ho, ho, L.cout, device=self.device, dtype=torch.bfloat16)
        else:
            ho = (L.hin + 2 * L.pad - L.r) // L.stride + 1
            if L.gap:
                out = torch.empty(n, L.cout, device=self.device, dtype=torch.float32)
            else:
                out = torch.empty(n, ho, ho, L.cout, device=self.device, dtype=torch.bfloat16)
        rp = None
        if residual is not None:
            self._check_dev(residual, "residual")
            rp = residual.data_ptr()
        x2p = None
        if x2 is not None:
            self._check_dev(x2, "x2")
            x2p = x2.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.phdfx_run_layer2(self._h, layer_id, x.data_ptr(), x2p, rp, out.data_ptr(), n,
                                                  self._stream()), self._h)
        return out


    # ---- fused-span hook ---------------------------------------------------------------------------------------------
    def chain_span(self, layer_id: int) -> int:
        """Number of execution-list entries the fused launch starting at `layer_id` covers (0 = none starts there)."""
        return int(self._lib.phdfx_chain_span(self._h, layer_id))

    @torch.no_grad()
    def run_chain(self, first_layer_id: int, t1: torch.Tensor, x_or_res: torch.Tensor):
        """Run the fused bottleneck chain that starts at conv2 = `first_layer_id` on explicit NHWC bf16 tensors:
        t1 [n,H,H,width], x_or_res = down-sample source [n,H,H,64] or identity residual [n,H,H,4*width]
        (H, width = 56, 64 in layer1; 28, 128 in layer2).
        Returns (out [n,H,H,4*width], next block's t1 [n,H,H,cout] or None)."""
        span = self.chain_span(first_layer_id)
        if span == 0:
            raise RuntimeError(f"no fused chain starts at layer {first_layer_id}")
        self._check_dev(t1, "t1")
        self._check_dev(x_or_res, "x_or_res")
        n = t1.shape[0]
        L3 = self.plan.layers[first_layer_id + 1]
        hw = L3.hin
        if tuple(t1.shape[1:]) != (hw, hw, L3.cin):
            raise RuntimeError(f"t1 must be [n,{hw},{hw},{L3.cin}], got {tuple(t1.shape)}")
        want_c = L3.cin2 if L3.in2_buf >= 0 else L3.cout
        if tuple(x_or_res.shape) != (n, hw, hw, want_c):
            raise RuntimeError(f"x_or_res must be [{n},{hw},{hw},{want_c}], got {tuple(x_or_res.shape)}")
        out = torch.empty(n, hw, hw, L3.cout, device=self.device, dtype=torch.bfloat16)
        t1n = None
        if span == 3:
            t1n = torch.empty(n, hw, hw, self.plan.layers[first_layer_id + 2].cout, device=self.device,
                              dtype=torch.bfloat16)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.phdfx_run_chain(self._h, first_layer_id, t1.data_ptr(), x_or_res.data_ptr(),
                                                 out.data_ptr(), t1n.data_ptr() if t1n is not None else None, n,
                                                 self._stream()), self._h)
        return out, t1n


class ExtractGraph:
    """One whole step of the hot path (K1 + fused stem/max-pool + 52 convs) frozen into a CUDA graph.

    The tensors given at capture time are the graph's fixed inputs / output: refill `frames` (and `boxes`) in place,
    call replay(), read `out`.  Replaying is bit-identical to calling extract_u8 and removes the per-kernel launch
    overhead (~10 us of fixed cost per kernel is what bounds small batches)."""

    def __init__(self, eng: B200Backbone, frames: torch.Tensor, boxes: Optional[torch.Tensor], flip_w: bool,
                 out: Optional[torch.Tensor]):
        n = frames.shape[0]
        if n > eng.max_frames:
            raise RuntimeError(f"a graph holds one engine call: n = {n} > max_frames = {eng.max_frames}")
        self.eng, self.frames, self.boxes, self.flip_w = eng, frames, boxes, flip_w
        self.out = out if out is not None else torch.empty(n, eng.FEAT_DIM, device=eng.device, dtype=torch.float32)
        side = torch.cuda.Stream(eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):  # warm-up outside capture (sets function attributes, builds nothing lazily)
            eng.extract_u8(frames, boxes, flip_w=flip_w, out=self.out)
        torch.cuda.current_stream(eng.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: other host threads (a DataLoader's pin-memory thread calling cudaHostAlloc, a shard writer) may
        # make CUDA calls while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            eng.extract_u8(frames, boxes, flip_w=flip_w, out=self.out)
        self.launches = eng.launches

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.out


class CapturedCall:
    """fn() — engine calls on fixed tensors — frozen into a CUDA graph (see ExtractGraph)."""

    def __init__(self, eng: B200Backbone, fn):
        side = torch.cuda.Stream(eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream(eng.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.result = fn()
        self.launches = eng.last_launch_count  # kernels of the last engine call inside fn()

    def replay(self):
        self.graph.replay()
        return self.result
