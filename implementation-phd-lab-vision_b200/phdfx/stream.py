"""Host-buffer front door: pinned, double-buffered H2D -> extract_u8 -> D2H pipeline.

Replaces the reference's per-variant `v_video.to(device)` ... `feats.to(dtype).cpu()` round trip
(src/preprocess_resnet_features.py:288-297), whose `.cpu()` is a blocking copy into pageable memory that serialises
host and GPU every variant.  Here the upload of batch i+1 and the download of batch i-1 overlap the trunk of batch i
on separate streams; the only synchronisation is at the end of the call.
"""
from __future__ import annotations

from typing import Optional

import torch

from .backbone import B200Backbone


class StreamingExtractor:
    """features = StreamingExtractor(engine)(frames_host_u8 [N,H,W,3], boxes_host [N,4] | None) -> host fp32 [N,2048]"""

    def __init__(self, engine: B200Backbone, batch: int = 256, save_fp16: bool = False, use_graphs: bool = True):
        if batch > engine.max_frames:
            raise RuntimeError(f"batch {batch} > engine.max_frames {engine.max_frames}")
        self.eng = engine
        self.batch = int(batch)
        self.dev = engine.device
        self.out_dtype = torch.float16 if save_fp16 else torch.float32  # --save-fp16 (:146,:285)
        self.copy_in = torch.cuda.Stream(self.dev)
        self.compute = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)
        self.use_graphs = use_graphs
        self._graphs = {}  # (slot, with_boxes, flip_w) -> ExtractGraph over the slot's fixed device buffers
        self._shape = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.launches = 0

    def _alloc(self, H: int, W: int):
        if self._shape == (H, W):
            return
        self._shape = (H, W)
        self._graphs = {}
        self.d_frames = [torch.empty(self.batch, H, W, 3, dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.d_boxes = [torch.empty(self.batch, 4, dtype=torch.int32, device=self.dev) for _ in range(2)]
        self.d_feats = [torch.empty(self.batch, self.eng.FEAT_DIM, dtype=torch.float32, device=self.dev)
                        for _ in range(2)]
        self.d_out = [torch.empty(self.batch, self.eng.FEAT_DIM, dtype=self.out_dtype, device=self.dev)
                      for _ in range(2)]
        self.ev_in = [torch.cuda.Event() for _ in range(2)]      # upload of slot done
        self.ev_done = [torch.cuda.Event() for _ in range(2)]    # compute of slot done
        self.ev_out = [torch.cuda.Event() for _ in range(2)]     # download of slot done

    @torch.no_grad()
    def __call__(self, frames: torch.Tensor, boxes: Optional[torch.Tensor] = None, flip_w: bool = False,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if frames.device.type != "cpu" or frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise RuntimeError("frames must be a host uint8 tensor [N,H,W,3]")
        n, H, W, _ = frames.shape
        if not frames.is_pinned():
            frames = frames.pin_memory()
        if boxes is not None:
            boxes = boxes.to(torch.int32)
            if tuple(boxes.shape) != (n, 4):
                raise RuntimeError("boxes must be [N,4] (top, left, h, w)")
            bt, bl, bh, bw = boxes.unbind(1)
            if bool(((bt < 0) | (bl < 0) | (bh < 1) | (bw < 1) | (bt + bh > H) | (bl + bw > W)).any()):
                raise RuntimeError(f"boxes must lie inside the {H}x{W} frames (top, left >= 0; h, w >= 1)")
            if not boxes.is_pinned():
                boxes = boxes.pin_memory()
        if out is None:
            out = torch.empty(n, self.eng.FEAT_DIM, dtype=self.out_dtype).pin_memory()
        self._alloc(H, W)
        self.h2d_bytes = self.d2h_bytes = self.launches = 0
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.copy_in, self.compute, self.copy_out):
            s.wait_stream(cur)
        for i, lo in enumerate(range(0, n, self.batch)):
            m = min(self.batch, n - lo)
            slot = i & 1
            with torch.cuda.stream(self.copy_in):
                if i >= 2:
                    self.copy_in.wait_event(self.ev_done[slot])  # slot's previous compute has consumed its input
                self.d_frames[slot][:m].copy_(frames[lo:lo + m], non_blocking=True)
                self.h2d_bytes += m * H * W * 3
                if boxes is not None:
                    self.d_boxes[slot][:m].copy_(boxes[lo:lo + m], non_blocking=True)
                    self.h2d_bytes += m * 16
                self.ev_in[slot].record(self.copy_in)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self.ev_in[slot])
                if i >= 2:
                    self.compute.wait_event(self.ev_out[slot])  # slot's previous result has been downloaded
                if self.use_graphs and m == self.batch:
                    key = (slot, boxes is not None, bool(flip_w))
                    g = self._graphs.get(key)
                    if g is None:  # first full batch on this slot: capture (the capture run does the work too)
                        g = self.eng.capture_extract(self.d_frames[slot], self.d_boxes[slot] if boxes is not None
                                                     else None, flip_w=flip_w, out=self.d_feats[slot])
                        self._graphs[key] = g
                    g.replay()
                    self.launches += g.launches
                else:
                    self.eng.extract_u8(self.d_frames[slot][:m],
                                        self.d_boxes[slot][:m] if boxes is not None else None, flip_w=flip_w,
                                        out=self.d_feats[slot][:m])
                    self.launches += self.eng.launches
                if self.out_dtype != torch.float32:
                    self.d_out[slot][:m].copy_(self.d_feats[slot][:m])
                self.ev_done[slot].record(self.compute)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(self.ev_done[slot])
                src = self.d_feats[slot] if self.out_dtype == torch.float32 else self.d_out[slot]
                out[lo:lo + m].copy_(src[:m], non_blocking=True)
                self.d2h_bytes += m * self.eng.FEAT_DIM * out.element_size()
                self.ev_out[slot].record(self.copy_out)
        for s in (self.copy_in, self.compute, self.copy_out):
            cur.wait_stream(s)
        torch.cuda.current_stream(self.dev).synchronize()
        return out
