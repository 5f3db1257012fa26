"""BN folding, weight packing and the execution list for libphdfx.

Input is the module the reference builds (src/preprocess_resnet_features.py:207-209):
    backbone = nn.Sequential(*list(torchvision.models.resnet50(...).children())[:-1]).eval()
i.e. [conv1, bn1, relu, maxpool, layer1, layer2, layer3, layer4, avgpool]; a full torchvision ResNet is accepted too.

Every conv + eval-mode BatchNorm pair (torchvision models/resnet.py:134-138,198,242) is folded in fp32:
    w' = w * gamma / sqrt(var + eps)        b' = beta - mean * gamma / sqrt(var + eps)
and w' is packed K-major as [Cout][R][S][Cin] bf16 — the K order the im2col tiles use.  The stem is packed as
[7 (filter row)][64 (cout)][32 (k = (s+1)*4 + c)], matching the 8-pixel x 4-channel window of the NHWC4p input.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.nn as nn

from ._lib import PHDFX_CONV, PHDFX_MAXPOOL, PHDFX_STEM, PHDFX_STEM_POOL, LayerDesc

_ALIGN = 64  # elements; keeps every layer's weight block 128-byte aligned for TMA


@dataclass
class Plan:
    weights: torch.Tensor  # bf16 [n_weights], CPU, contiguous
    bias: torch.Tensor  # fp32 [n_bias], CPU, contiguous
    layers: List[LayerDesc]
    names: List[str]  # human-readable layer names, same order as `layers`
    block_first: List[int] = None  # execution-list index of every bottleneck block's conv1 (16 entries)


def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """Return (w', b') in fp32: eval-mode BN folded into the conv (exact in fp32)."""
    if conv.bias is not None:
        raise ValueError("ResNet convs have no bias")
    w = conv.weight.detach().to(torch.float32).cpu()
    gamma = bn.weight.detach().to(torch.float32).cpu()
    beta = bn.bias.detach().to(torch.float32).cpu()
    mean = bn.running_mean.detach().to(torch.float32).cpu()
    var = bn.running_var.detach().to(torch.float32).cpu()
    scale = gamma / torch.sqrt(var + bn.eps)
    return w * scale[:, None, None, None], beta - mean * scale


def pack_conv(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, R, S] fp32 -> flat bf16 [Cout][R][S][Cin]."""
    return w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).reshape(-1)


def pack_stem(w: torch.Tensor) -> torch.Tensor:
    """[64, 3, 7, 7] fp32 -> flat bf16 [7][64][32] with k = (s+1)*4 + c (k = 0..3 and channel 3 are zero)."""
    cout, cin, r, s = w.shape
    assert (cout, cin, r, s) == (64, 3, 7, 7), w.shape
    out = torch.zeros(7, 64, 8, 4, dtype=torch.float32)
    # out[r, co, s+1, c] = w[co, c, r, s]
    out[:, :, 1:8, 0:3] = w.permute(2, 0, 3, 1)
    return out.to(torch.bfloat16).reshape(-1)


def pack_stem_pool(w: torch.Tensor) -> torch.Tensor:
    """[64, 3, 7, 7] fp32 -> flat bf16: the shared-memory image of the fused stem kernel's B operands (K-major, no swizzle:
    8x16 B core matrices), k = chunk*8 + e = (s+1)*4 + c:
      [7 (filter row)][4 (k-chunk)][64 (cout)][8]                      one filter row, N = 64
      [5 (e = 2..6)][4 (k-chunk)][128 = filter row e | filter row e-2][8]  stacked pairs, N = 128: an input row that is
          filter row e of an even conv row is filter row e-2 of the next (odd) one (stem_pool_sm100.cuh, MMA warp)."""
    cout, cin, r, s = w.shape
    assert (cout, cin, r, s) == (64, 3, 7, 7), w.shape
    out = torch.zeros(7, 64, 8, 4, dtype=torch.float32)
    out[:, :, 1:8, 0:3] = w.permute(2, 0, 3, 1)
    rows = out.reshape(7, 64, 4, 8).permute(0, 2, 1, 3).contiguous()  # [7][4][64][8]
    pairs = torch.cat([rows[2:7], rows[0:5]], dim=2)  # [5][4][128][8]: rows 0..63 = filter row e, 64..127 = e-2
    return torch.cat([rows.reshape(-1), pairs.contiguous().reshape(-1)]).to(torch.bfloat16)


def _children(backbone: nn.Module):
    """(conv1, bn1, maxpool, [layer1..4]) from the reference's Sequential or a torchvision ResNet."""
    if isinstance(backbone, nn.Sequential):
        mods = list(backbone.children())
        if len(mods) < 8:
            raise ValueError("expected nn.Sequential(*list(resnet.children())[:-1])")
        return mods[0], mods[1], mods[3], mods[4:8]
    return backbone.conv1, backbone.bn1, backbone.maxpool, [backbone.layer1, backbone.layer2, backbone.layer3,
                                                            backbone.layer4]


def build_plan(backbone: nn.Module, fuse_stem_pool: bool = True, fuse_downsample: bool = True,
               stage_after_blocks=(3, 7, 13)) -> Plan:
    """fuse_stem_pool=True (default): conv1+bn1+relu+maxpool is ONE launch (stem_pool_sm100.cuh).  False keeps the
    separate implicit-GEMM stem and max-pool kernels (used for A/B measurements and their own parity tests).
    fuse_downsample=True (default): in the first block of each stage the down-sample branch is accumulated into
    conv3's GEMM (K concatenated, weights [cout][width + cin], bias b3 + bd) instead of being a launch and a tensor
    of its own:  out = relu(conv3(t2) + downsample(x))  (resnet.py:154-161).
    stage_after_blocks: bottleneck blocks (1-based, in network order; default = the ends of layer1, layer2, layer3)
    after which the execution schedule may be cut into frame-wave stages (phdfx_set_schedule): the tensors that cross
    such a cut — the block's output and the next block's conv1 output, which a fused chain launch produces on the
    near side of the cut — get arena buffer ids of their own (6, 7, ...), so every other buffer of a stage can be
    wave-local scratch (PHDFX_SCHED_REUSE)."""
    conv1, bn1, maxpool, stages = _children(backbone)
    cuts = set(int(b) for b in stage_after_blocks)
    next_id = 6
    carry_t1 = None
    block_first = []
    chunks, biases, layers, names = [], [], [], []
    w_cursor = 0
    b_cursor = 0

    def add_weights(packed: torch.Tensor, bias: torch.Tensor):
        nonlocal w_cursor, b_cursor
        w_off, b_off = w_cursor, b_cursor
        pad = (-packed.numel()) % _ALIGN
        chunks.append(packed)
        if pad:
            chunks.append(torch.zeros(pad, dtype=torch.bfloat16))
        w_cursor += packed.numel() + pad
        biases.append(bias.to(torch.float32).reshape(-1))
        b_cursor += bias.numel()
        return w_off, b_off

    def desc(**kw) -> LayerDesc:
        d = LayerDesc()
        base = dict(kind=PHDFX_CONV, cin=0, cout=0, r=1, s=1, stride=1, pad=0, hin=0, win=0, relu=0, in_buf=0,
                    out_buf=0, res_buf=-1, gap=0, in2_buf=-1, cin2=0, stride2=1, hin2=0, w_off=0, b_off=0)
        base.update(kw)
        for k, v in base.items():
            setattr(d, k, int(v))
        return d

    # stem: conv1 + bn1 + relu (resnet.py:268-270), then maxpool (:271)
    w, b = fold_conv_bn(conv1, bn1)
    if fuse_stem_pool:
        w_off, b_off = add_weights(pack_stem_pool(w), b)
        layers.append(desc(kind=PHDFX_STEM_POOL, cin=3, cout=64, r=7, s=7, stride=2, pad=3, hin=224, win=224,
                           relu=1, in_buf=0, out_buf=2, w_off=w_off, b_off=b_off))
        names.append("conv1+maxpool")
    else:
        w_off, b_off = add_weights(pack_stem(w), b)
        layers.append(desc(kind=PHDFX_STEM, cin=3, cout=64, r=7, s=7, stride=2, pad=3, hin=224, win=224, relu=1,
                           in_buf=0, out_buf=1, w_off=w_off, b_off=b_off))
        names.append("conv1")
        layers.append(desc(kind=PHDFX_MAXPOOL, cin=64, cout=64, r=3, s=3, stride=2, pad=1, hin=112, win=112,
                           in_buf=1, out_buf=2))
        names.append("maxpool")

    # block input / block output ping-pong on 2 / 1; conv1 out (t1) and conv2 out (t2) alternate between 3 and 4 from
    # block to block, so a block's t1 and the NEXT block's t1 never share a buffer — the fused layer1 chain
    # (bottleneck_chain_sm100.cuh) reads the former (with halo rows of neighbouring tiles) while writing the latter;
    # 5 = down-sample out (unfused plan only)
    x_buf, o_buf = 2, 1
    h = 56
    n_blocks = sum(len(s) for s in stages)
    blk_i = 0
    for li, stage in enumerate(stages):
        for bi, blk in enumerate(stage):
            blk_i += 1
            last = blk_i == n_blocks
            name = f"layer{li + 1}.{bi}"
            cin = blk.conv1.in_channels
            width = blk.conv1.out_channels
            cout = blk.conv3.out_channels
            stride = blk.conv2.stride[0]
            t1_buf, t2_buf = (3, 4) if blk_i % 2 == 1 else (4, 3)
            if carry_t1 is not None:  # first block behind a cut: its conv1 output crosses the cut inside a chain launch
                t1_buf, carry_t1 = carry_t1, None
            out_id = o_buf
            if blk_i in cuts and not last:
                if next_id + 1 > 15:
                    raise ValueError("too many stage cuts for the 16 arena buffer ids")
                out_id, carry_t1, next_id = next_id, next_id + 1, next_id + 2
            block_first.append(len(layers))
            # conv1 1x1 + bn1 + relu (resnet.py:146-148)
            w, b = fold_conv_bn(blk.conv1, blk.bn1)
            w_off, b_off = add_weights(pack_conv(w), b)
            layers.append(desc(cin=cin, cout=width, hin=h, win=h, relu=1, in_buf=x_buf, out_buf=t1_buf, w_off=w_off,
                               b_off=b_off))
            names.append(name + ".conv1")
            # conv2 3x3 (stride lives here: v1.5, resnet.py:109-113) + bn2 + relu (:150-152)
            w, b = fold_conv_bn(blk.conv2, blk.bn2)
            w_off, b_off = add_weights(pack_conv(w), b)
            layers.append(desc(cin=width, cout=width, r=3, s=3, stride=stride, pad=1, hin=h, win=h, relu=1,
                               in_buf=t1_buf, out_buf=t2_buf, w_off=w_off, b_off=b_off))
            names.append(name + ".conv2")
            ho = (h + 2 - 3) // stride + 1
            res_buf = x_buf
            w3, b3 = fold_conv_bn(blk.conv3, blk.bn3)
            if blk.downsample is not None and fuse_downsample:
                # conv3 and the down-sample branch write the same pixels: one GEMM over K = width + cin
                dconv, dbn = blk.downsample[0], blk.downsample[1]
                wd, bd = fold_conv_bn(dconv, dbn)
                wcat = torch.cat([w3.reshape(cout, width), wd.reshape(cout, cin)], dim=1)  # [cout][width | cin]
                w_off, b_off = add_weights(wcat.contiguous().to(torch.bfloat16).reshape(-1), b3 + bd)
                layers.append(desc(cin=width, cout=cout, hin=ho, win=ho, relu=1, in_buf=t2_buf, out_buf=out_id, res_buf=-1,
                                   gap=1 if last else 0, in2_buf=x_buf, cin2=cin, stride2=dconv.stride[0], hin2=h,
                                   w_off=w_off, b_off=b_off))
                names.append(name + ".conv3+downsample")
            else:
                if blk.downsample is not None:
                    # downsample 1x1/stride + bn, no relu (resnet.py:157-158, 239-243)
                    dconv, dbn = blk.downsample[0], blk.downsample[1]
                    w, b = fold_conv_bn(dconv, dbn)
                    w_off, b_off = add_weights(pack_conv(w), b)
                    layers.append(desc(cin=cin, cout=cout, stride=dconv.stride[0], hin=h, win=h, relu=0,
                                       in_buf=x_buf, out_buf=5, w_off=w_off, b_off=b_off))
                    names.append(name + ".downsample")
                    res_buf = 5
                # conv3 1x1 + bn3 + residual + relu (:154-161); the last one also fuses avgpool (:278)
                w_off, b_off = add_weights(pack_conv(w3), b3)
                layers.append(desc(cin=width, cout=cout, hin=ho, win=ho, relu=1, in_buf=t2_buf, out_buf=out_id,
                                   res_buf=res_buf, gap=1 if last else 0, w_off=w_off, b_off=b_off))
                names.append(name + ".conv3")
            # the block's output is the next block's input; the next output goes to a scratch id (1 / 2) it is not
            x_buf, o_buf = out_id, (2 if out_id == 1 else 1)
            h = ho
    return Plan(weights=torch.cat(chunks).contiguous(), bias=torch.cat(biases).contiguous(), layers=layers,
                names=names, block_first=block_first)


def randomize_bn_(backbone: nn.Module, seed: int = 1) -> nn.Module:
    """Give every BatchNorm non-trivial running stats / affine params (seeded).  Random-init BN is the identity,
    which would leave BN folding untested; magnitudes keep activations O(1)."""
    g = torch.Generator().manual_seed(seed)
    for m in backbone.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            with torch.no_grad():
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(0.5 + torch.rand(n, generator=g))
                m.weight.copy_(0.6 + 0.5 * torch.rand(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))
    return backbone
