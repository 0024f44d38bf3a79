"""FPN4 (mvs4net_utils.py:419-502) on libmvster_b200: all views of a frame in one batch, channels-last,
BN folded, the 3x3 stride-1 layers with Cin >= 16 (73 % of the FLOPs) optionally on the tcgen05 kernel.

FPN4 is outside the named hot path but feeds it (SURVEY.md 8f "next" #1); through cuDNN fp32 it was
75 % of the step.  Layout: every intermediate is NHWC fp32, so the four outputs are directly the
feature tensors the warp kernel consumes.
"""
from __future__ import annotations

import os
from typing import Dict, Mapping, Optional, Sequence

import torch

from . import _lib, capi, packing

Tensor = torch.Tensor


def _fold2d(sd: Mapping[str, Tensor], p: str):
    """Conv2d(no bias) + BN(eval) -> ([k*k][Cin][Cout], [Cout])."""
    w = sd[p + ".conv.weight"].detach().cpu().double()
    s, t = packing.bn_scale_shift({k: v.detach().cpu() for k, v in sd.items() if k.startswith(p + ".bn")}, p + ".bn")
    wf = w * s.view(-1, 1, 1, 1)
    k = w.shape[2]
    return wf.permute(2, 3, 1, 0).reshape(k * k, w.shape[1], w.shape[0]).float().contiguous(), t.float().contiguous()


def _plain2d(w: Tensor):
    w = w.detach().cpu()
    k = w.shape[2]
    return w.permute(2, 3, 1, 0).reshape(k * k, w.shape[1], w.shape[0]).float().contiguous()


def pack_fpn(sd: Mapping[str, Tensor], prefix: str = "feature", out_perm: Optional[Mapping[int, Sequence[int]]] = None) -> Dict[str, Tensor]:
    """All FPN4 weights in kernel layouts (CPU tensors): '<layer>.w' [taps][Cin][Cout], '<layer>.b' [Cout],
    '<layer>.tc' K-major [hi|lo] slabs for the 3x3 stride-1 layers with Cin >= 16.
    ``out_perm[i]`` (i = 1..4): output-channel order of ``out{i}`` - memory position p of the stage-i features holds channel
    ``out_perm[i][p]`` (the group-interleaved layout of MVSTER_ET_INTERLEAVED, capi.interleave_perm); every packed form of
    that convolution, the fused last level included, is built from the permuted weights, so the layout costs nothing."""
    out: Dict[str, Tensor] = {}
    p = prefix
    for name in ("conv0.0", "conv0.1", "conv1.0", "conv1.1", "conv1.2", "conv2.0", "conv2.1", "conv2.2", "conv3.0", "conv3.1", "conv3.2"):
        w, b = _fold2d(sd, f"{p}.{name}")
        out[name + ".w"], out[name + ".b"] = w, b
        if w.shape[0] == 9 and w.shape[1] >= 16:
            out[name + ".tc"] = packing.pack_tc2_weights(w, 3)
        if w.shape[1] >= 8:  # generation-3 kernel: every layer after the 3-channel stem, stride-2 5x5 layers included
            kk, ss = (5 if w.shape[0] == 25 else 3), (2 if name.endswith(".0") else 1)
            out[name + ".tc3"] = packing.pack_tc3_weights(w, 1, kk, ss)
            out[name + ".tc3h"] = packing.pack_tc3_weights(w, 1, kk, ss, split=2)  # two fp16 terms (MVSTER_TC3_FP16X2)
    for i in (1, 2, 3):
        w = sd[f"{p}.inner{i}.weight"].detach().cpu()
        out[f"inner{i}.w"] = w.reshape(w.shape[0], w.shape[1]).t().float().contiguous()  # [Clat][64]
        out[f"inner{i}.b"] = sd[f"{p}.inner{i}.bias"].detach().cpu().float().contiguous()
    for i in (1, 2, 3, 4):
        w = _plain2d(sd[f"{p}.out{i}.weight"])
        if out_perm and out_perm.get(i) is not None and list(out_perm[i]) != list(range(w.shape[2])):
            w = w[:, :, list(out_perm[i])].contiguous()
        out[f"out{i}.w"] = w
        if w.shape[0] == 9:
            out[f"out{i}.tc"] = packing.pack_tc2_weights(w, 3)
        if w.shape[2] >= 8 and i < 4:
            out[f"out{i}.tc3"] = packing.pack_tc3_weights(w, 1, 3 if w.shape[0] == 9 else 1, 1)
            out[f"out{i}.tc3h"] = packing.pack_tc3_weights(w, 1, 3 if w.shape[0] == 9 else 1, 1, split=2)
    # fused last level (fpn.cu: fpn_out4_gather_kernel): out4(up2(top2) + inner3(c0)) with the 3x3 conv's channel
    # mixing moved in front of the up-sampling - composite weights, all built in fp64
    w4 = out["out4.w"].double()                       # [9][64][8]  (tap, top channel m, out channel o)
    wi3, bi3 = out["inner3.w"].double(), out["inner3.b"].double()  # [8][64], [64]
    out["out4.wc"] = torch.einsum("cm,tmo->tco", wi3, w4).float().contiguous()   # [9][8][8]
    out["out4.bt"] = torch.einsum("m,tmo->to", bi3, w4).float().contiguous()     # [9][8]
    u_w = w4.permute(1, 0, 2).reshape(1, 64, 72).float().contiguous()            # U channel = tap*8 + o
    out["out4.u_w"] = u_w
    out["out4.u_tc"] = packing.pack_tc2_weights(u_w, 3)
    # generation 3: one 64 -> 72 point-wise GEMM (N padded to 80) that writes each tap's 8 channels as its own plane
    out["out4.u_tc3"] = packing.pack_tc3_weights(u_w, 1, 1, 1)
    out["out4.u_tc3h"] = packing.pack_tc3_weights(u_w, 1, 1, 1, split=2)
    return out


def _conv2d(x: Tensor, w: Tensor, b: Optional[Tensor], k: int, stride: int, relu: bool) -> Tensor:
    N, H, W, Cin = x.shape
    cout = w.shape[2]
    y = torch.empty((N, (H - 1) // stride + 1, (W - 1) // stride + 1, cout), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_conv2d_nhwc_f32(capi._ptr(x), capi._ptr(w), capi._ptr(b), capi._ptr(y), N, H, W, Cin, cout, k, stride,
                                                  int(relu), capi._stream()), "mvster_conv2d_nhwc_f32")
    return y


def _conv_tc3(x: Tensor, wts: Dict[str, Tensor], name: str, k: int, stride: int, relu: bool, split: int = 3) -> Tensor:
    """any FPN conv after the stem on the persistent tcgen05 kernel (conv_tc3.cu); split 3 = 3 x bf16, 2 = 2 x fp16 operands."""
    N, H, W, Cin = x.shape
    cout = wts[name + ".w"].shape[2]
    y = capi.conv_tc3(x.view(N, 1, H, W, Cin), wts[name + (".tc3h" if split == 2 else ".tc3")], wts.get(name + ".b"), cout, 1, k, stride,
                      relu, split=split)
    return y.view(N, y.shape[2], y.shape[3], cout)


def _conv3x3(x: Tensor, wts: Dict[str, Tensor], name: str, relu: bool, npass: int, gen: int = 2) -> Tensor:
    """3x3 stride-1 layer: tensor cores (gen 3: 3 x bf16; gen 2: npass 3 = 3xTF32, 1 = TF32) or CUDA cores (npass 0)."""
    if npass and gen == 3 and (name + ".tc3") in wts:
        return _conv_tc3(x, wts, name, 3, 1, relu, 2 if npass == 2 else 3)
    w, b = wts[name + ".w"], wts.get(name + ".b")
    if npass == 2:
        npass = 3  # a layer without generation-3 weights (the literal out4): 3xTF32 on the generation-2 kernel
    if npass and (name + ".tc") in wts:
        N, H, W, Cin = x.shape
        tc = wts[name + ".tc"]
        half = tc.numel() // 2
        y = capi.conv3d_tc2(x.view(N, 1, H, W, Cin), tc if npass == 3 else tc[:half].contiguous(), b, w.shape[2], 1, relu, npass=npass)
        return y.view(N, H, W, w.shape[2])
    return _conv2d(x, w, b, 3, 1, relu)


def _merge(top: Tensor, lat: Tensor, w: Tensor, b: Tensor) -> Tensor:
    N, H, W, CL = lat.shape
    out = torch.empty((N, H, W, 64), device=lat.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_fpn_merge_f32(capi._ptr(top), capi._ptr(lat), capi._ptr(w), capi._ptr(b), capi._ptr(out), N, H, W, CL,
                                                capi._stream()), "mvster_fpn_merge_f32")
    return out


def run_fpn(wts: Dict[str, Tensor], imgs: Tensor, npass: int = 0, fused_last: bool = True, gen: int = 2,
            on_level=None) -> Dict[str, Tensor]:
    """imgs [N,3,H,W] contiguous NCHW fp32 on the GPU -> {'stage1'..'stage4': [N,h,w,C] NHWC}.
    npass: 0 = every layer on the CUDA cores (exact fp32), 3 / 1 = 3x3 stride-1 layers with Cin >= 16 on tcgen05
    (3xTF32 / TF32); gen = 3 (with npass = 3): every conv after the 3-channel stem, the stride-2 5x5 layers and the 1x1 out1
    included, on the persistent 3 x bf16 kernel.  fused_last: algebraically fused last level (never forms the 64-channel
    full-res map).  on_level(k, tensor) is called right after the launches producing 'stage{k+1}' are enqueued (coarsest
    first), so a caller can record an event and start that cascade stage on another stream."""
    capi._chk(imgs, "imgs")
    N, three, H, W = imgs.shape
    if three != 3 or H % 8 or W % 8:
        raise ValueError(f"imgs must be [N,3,H,W] with H,W multiples of 8, got {tuple(imgs.shape)}")
    c0 = torch.empty((N, H, W, 8), device=imgs.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_conv_first_f32(capi._ptr(imgs), capi._ptr(wts["conv0.0.w"]), capi._ptr(wts["conv0.0.b"]), capi._ptr(c0),
                                                 N, H, W, capi._stream()), "mvster_conv_first_f32")
    g3 = bool(npass) and gen == 3
    sp = 2 if npass == 2 else 3  # generation 3 only: npass = 2 selects the two-fp16-term arithmetic
    # conv0.1 (8 -> 8 at full resolution): on the tensor cores two taps share each MMA (K = 2 x 8 channels); MVSTER_FPN_C01=simt
    # keeps it on the CUDA cores (107 us at cfg2; 124 us on the tensor cores before the tap pairing)
    if g3 and os.environ.get("MVSTER_FPN_C01", "tc") != "simt":
        c0 = _conv_tc3(c0, wts, "conv0.1", 3, 1, True, sp)
    else:
        c0 = _conv2d(c0, wts["conv0.1.w"], wts["conv0.1.b"], 3, 1, True)
    levels = [c0]
    x = c0
    for L in (1, 2, 3):
        x = _conv_tc3(x, wts, f"conv{L}.0", 5, 2, True, sp) if g3 else _conv2d(x, wts[f"conv{L}.0.w"], wts[f"conv{L}.0.b"], 5, 2, True)
        x = _conv3x3(x, wts, f"conv{L}.1", True, npass, gen)
        x = _conv3x3(x, wts, f"conv{L}.2", True, npass, gen)
        levels.append(x)
    c0, c1, c2, c3 = levels
    notify = on_level if on_level is not None else (lambda k, t: None)
    out = {"stage1": _conv_tc3(c3, wts, "out1", 1, 1, False, sp) if g3 else _conv2d(c3, wts["out1.w"], None, 1, 1, False)}
    notify(0, out["stage1"])
    top = _merge(c3, c2, wts["inner1.w"], wts["inner1.b"])
    out["stage2"] = _conv3x3(top, wts, "out2", False, npass, gen)
    notify(1, out["stage2"])
    top = _merge(top, c1, wts["inner2.w"], wts["inner2.b"])
    out["stage3"] = _conv3x3(top, wts, "out3", False, npass, gen)
    notify(2, out["stage3"])
    if fused_last:
        out["stage4"] = _fused_last_level(wts, top, c0, npass, gen)
    else:  # literal form: materialise the 64-channel full-resolution map, then the 3x3 conv
        top = _merge(top, c0, wts["inner3.w"], wts["inner3.b"])
        out["stage4"] = _conv3x3(top, wts, "out4", False, npass, gen)
    notify(3, out["stage4"])
    return out


def _fused_last_level(wts: Dict[str, Tensor], top2: Tensor, c0: Tensor, npass: int, gen: int = 2) -> Tensor:
    """stage4 = out4(up2(top2) + inner3(c0)) = sum_tap [ up2(W4[tap].top2) + (W4[tap] Wi3).c0 + W4[tap].bi3 ](p + tap)."""
    N, h, w, _ = top2.shape
    lib = _lib.load()
    uc = 72
    if npass and gen == 3:  # U planar [9][N][h][w][8]: one point-wise GEMM on the persistent 3 x bf16 kernel
        U = torch.empty((9, N, h, w, 8), device=top2.device, dtype=torch.float32)
        h16 = npass == 2
        _lib.check(lib.mvster_pointwise_tc3_blocks_ex_f32(capi._ptr(top2), capi._ptr(wts["out4.u_tc3h" if h16 else "out4.u_tc3"]), capi._ptr(U),
                                                          N, h, w, 64, 72, 8, N * h * w * 8, capi.TC3_FP16X2 if h16 else 0, capi._stream()),
                   "mvster_pointwise_tc3_blocks_f32")
        uc = 8
    elif npass:
        tc = wts["out4.u_tc"]
        U = torch.empty((N, h, w, 72), device=top2.device, dtype=torch.float32)
        _lib.check(lib.mvster_pointwise_tc2_f32(capi._ptr(top2), capi._ptr(tc if npass == 3 else tc[:tc.numel() // 2].contiguous()), None,
                                                capi._ptr(U), N, h, w, 64, 72, 0, npass, capi._stream()), "mvster_pointwise_tc2_f32")
    else:
        U = _conv2d(top2, wts["out4.u_w"], None, 1, 1, False)
    out = torch.empty((N, 2 * h, 2 * w, 8), device=top2.device, dtype=torch.float32)
    _lib.check(lib.mvster_fpn_out4_gather_f32(capi._ptr(U), uc, capi._ptr(c0), capi._ptr(wts["out4.wc"]), capi._ptr(wts["out4.bt"]),
                                              capi._ptr(out), N, 2 * h, 2 * w, capi._stream()), "mvster_fpn_out4_gather_f32")
    return out
