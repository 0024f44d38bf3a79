"""Tensor-level wrappers over the C ABI: validate, take ``data_ptr()``s, launch on torch's
current stream.  Layout conventions are those of include/mvster_b200.h (channels-last
features, NDHWC cost volume).  CUDA tensors only - anything else raises."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib

Tensor = torch.Tensor

# flags of mvster_et_fuse_f32 (include/mvster_b200.h)
ET_PARTIAL = 1
ET_ACCUMULATE = 2
ET_GENERIC = 4
ET_NO_FUSE_D = 8
ET_SQDIFF = 16
ET_WINDOW = 32
ET_NO_WINDOW = 64
ET_TMA_ON = 128
ET_TMA_OFF = 256
ET_INTERLEAVED = 512
MAX_VIEWS = 16


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _chk(t: Tensor, name: str, shape: Optional[Sequence[int]] = None, dtype: torch.dtype = torch.float32) -> Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.MvsterLibraryError(f"{name}: a CUDA tensor is required (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: {dtype} required, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous, got strides {t.stride()} for shape {tuple(t.shape)}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def to_nhwc(x: Tensor) -> Tensor:
    """[B,C,H,W] (any strides) -> contiguous [B,H,W,C].  Free if x is already channels_last."""
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return v
    if x.is_contiguous() and x.dtype == torch.float32 and x.is_cuda:
        B, Cc, H, W = x.shape
        out = torch.empty((B, H, W, Cc), device=x.device, dtype=x.dtype)
        _lib.check(_lib.load().mvster_nchw_to_nhwc_f32(_ptr(x), _ptr(out), B, Cc, H, W, _stream()), "mvster_nchw_to_nhwc_f32")
        return out
    return v.contiguous()


def hypo_init_inverse(depth_values: Tensor, D: int, H: int, W: int, out: Optional[Tensor] = None) -> Tensor:
    dv = _chk(depth_values, "depth_values")
    B, n = dv.shape
    out = torch.empty((B, D, H, W), device=dv.device, dtype=torch.float32) if out is None else _chk(out, "hypo", (B, D, H, W))
    _lib.check(_lib.load().mvster_hypo_init_inverse_f32(_ptr(dv), n, _ptr(out), B, D, H, W, _stream()),
               "mvster_hypo_init_inverse_f32")
    return out


def hypo_schedule_inverse(inv_min: Tensor, inv_max: Tensor, D: int, H: int, W: int, out: Optional[Tensor] = None) -> Tensor:
    B = inv_min.shape[0]
    _chk(inv_min, "inv_min", (B, H // 2, W // 2))
    _chk(inv_max, "inv_max", (B, H // 2, W // 2))
    out = torch.empty((B, D, H, W), device=inv_min.device, dtype=torch.float32) if out is None else _chk(out, "hypo", (B, D, H, W))
    _lib.check(_lib.load().mvster_hypo_schedule_inverse_f32(_ptr(inv_min), _ptr(inv_max), _ptr(out), B, D, H, W, _stream()),
               "mvster_hypo_schedule_inverse_f32")
    return out


def pose(proj: Tensor, first_view: int = 1, n_views: Optional[int] = None) -> Tensor:
    """proj [B,Nv,2,4,4] -> [B,V,12] relative poses (R row-major, t) of source views
    first_view .. first_view+V-1 with respect to view 0."""
    _chk(proj, "proj")
    B, Nv = proj.shape[:2]
    if tuple(proj.shape[2:]) != (2, 4, 4):
        raise ValueError(f"proj: expected [B,Nv,2,4,4], got {tuple(proj.shape)}")
    V = Nv - first_view if n_views is None else n_views
    out = torch.empty((B, V, 12), device=proj.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_pose_f32(_ptr(proj), _ptr(out), B, Nv, first_view, V, _stream()), "mvster_pose_f32")
    return out


def et_fuse(ref: Tensor, srcs: Sequence[Tensor], pose: Tensor, hypo: Tensor, G: int, attn_temp: float,
            cost: Optional[Tensor] = None, wsum: Optional[Tensor] = None, partial: bool = False,
            accumulate: bool = False, generic: bool = False, group_cor: bool = True, fuse_d: bool = True,
            window: Optional[bool] = None, tma: Optional[bool] = None, interleaved: bool = False) -> Tensor:
    """ref [B,H,W,C], srcs V x [B,Hs,Ws,C], pose [B,V,12], hypo [B,D,H,W] -> cost [B,D,H,W,G].
    With ``partial`` the un-normalised accumulators are written to (cost, wsum).  ``group_cor=False``:
    per-channel squared difference, the cost volume then has C channels (pass G == C);
    ``fuse_d=False``: the reference's attn_fuse_d=False weighting.  ``window``: True / False force / forbid the
    window kernel (et_fuse_win.cuh); ``tma``: True / False force / forbid its TMA-staged form (et_fuse_tma.cuh: source boxes in
    shared memory); None leaves the choice to the library.  ``interleaved``: ref / srcs store each 8-channel block
    group-interleaved (MVSTER_ET_INTERLEAVED in include/mvster_b200.h; ``interleave_perm``)."""
    B, H, W, Cc = ref.shape
    _chk(ref, "ref")
    V = len(srcs)
    D = hypo.shape[1]
    _chk(hypo, "hypo", (B, D, H, W))
    Hs, Ws = srcs[0].shape[1:3]
    for i, s in enumerate(srcs):
        _chk(s, f"src[{i}]", (B, Hs, Ws, Cc))
    if cost is None:
        cost = torch.empty((B, D, H, W, G), device=ref.device, dtype=torch.float32)
    _chk(cost, "cost", (B, D, H, W, G))
    if (partial or accumulate) and wsum is None:
        wsum = torch.empty((B, D, H, W), device=ref.device, dtype=torch.float32)
    if wsum is not None:
        _chk(wsum, "wsum", (B, D, H, W))
    lib = _lib.load()
    for v0 in range(0, V, MAX_VIEWS):  # more than MAX_VIEWS views: chain launches through the partials
        chunk = srcs[v0:v0 + MAX_VIEWS]
        flags = ET_GENERIC if generic else 0
        if not fuse_d:
            flags |= ET_NO_FUSE_D
        if not group_cor:
            flags |= ET_SQDIFF
        if window is not None:
            flags |= ET_WINDOW if window else ET_NO_WINDOW
        if tma is not None:
            flags |= ET_TMA_ON if tma else ET_TMA_OFF
        if interleaved:
            flags |= ET_INTERLEAVED
        if partial or V > MAX_VIEWS:  # a chain stays un-normalised to its end: the division (with the 1e-8 seed) follows once
            flags |= ET_PARTIAL
        if accumulate or v0 > 0:
            flags |= ET_ACCUMULATE
        if (flags & 3) and wsum is None:
            wsum = torch.empty((B, D, H, W), device=ref.device, dtype=torch.float32)
        pose_c = _chk(pose[:, v0:v0 + len(chunk)].contiguous(), "pose", (B, len(chunk), 12))
        arr = (C.c_void_p * len(chunk))(*[s.data_ptr() for s in chunk])
        _lib.check(lib.mvster_et_fuse_f32(_ptr(ref), arr, len(chunk), _ptr(pose_c), _ptr(hypo), _ptr(cost), _ptr(wsum),
                                          B, Cc, G, D, H, W, Hs, Ws, float(attn_temp), flags, _stream()),
                   "mvster_et_fuse_f32")
    if V > MAX_VIEWS and not partial:
        et_normalize(cost, wsum)
    return cost


def cast_bf16(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """fp32 -> bf16 (round to nearest even) on the library's own kernel: the feature pyramid's outputs as the bf16-storage
    configuration stores them."""
    _chk(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if out is None else _chk(out, "out", tuple(x.shape), torch.bfloat16)
    _lib.check(_lib.load().mvster_cast_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "mvster_cast_bf16")
    return out


def et_fuse_bf16(ref: Tensor, srcs: Sequence[Tensor], pose: Tensor, hypo: Tensor, G: int, attn_temp: float,
                 window: Optional[bool] = None, interleaved: bool = False) -> Tensor:
    """bf16-storage form of ``et_fuse``: ref [B,H,W,C] and srcs V x [B,Hs,Ws,C] in torch.bfloat16 -> cost [B,D,H,W,G] in
    torch.bfloat16; fp32 arithmetic in between (mvster_et_fuse_bf16)."""
    B, H, W, Cc = ref.shape
    _chk(ref, "ref", dtype=torch.bfloat16)
    V = len(srcs)
    if V > MAX_VIEWS:
        raise ValueError(f"et_fuse_bf16: at most {MAX_VIEWS} source views per call (no partial sums with bf16 storage), got {V}")
    D = hypo.shape[1]
    _chk(hypo, "hypo", (B, D, H, W))
    Hs, Ws = srcs[0].shape[1:3]
    for i, s in enumerate(srcs):
        _chk(s, f"src[{i}]", (B, Hs, Ws, Cc), torch.bfloat16)
    _chk(pose, "pose", (B, V, 12))
    cost = torch.empty((B, D, H, W, G), device=ref.device, dtype=torch.bfloat16)
    flags = 0
    if window is not None:
        flags |= ET_WINDOW if window else ET_NO_WINDOW
    if interleaved:
        flags |= ET_INTERLEAVED
    arr = (C.c_void_p * V)(*[s.data_ptr() for s in srcs])
    _lib.check(_lib.load().mvster_et_fuse_bf16(_ptr(ref), arr, V, _ptr(pose), _ptr(hypo), _ptr(cost), B, Cc, G, D, H, W, Hs, Ws,
                                               float(attn_temp), flags, _stream()), "mvster_et_fuse_bf16")
    return cost


def interleave_perm(C: int, G: int) -> List[int]:
    """Channel stored at each memory position under MVSTER_ET_INTERLEAVED: ``stored[..., p] = natural[..., perm[p]]``."""
    cpg = C // G
    block = {2: [0, 2, 1, 3, 4, 6, 5, 7], 4: [0, 4, 1, 5, 2, 6, 3, 7]}.get(cpg)
    if block is None:
        return list(range(C))
    return [8 * (p // 8) + block[p % 8] for p in range(C)]


def et_last_kernel() -> str:
    """Kernel the last ``et_fuse`` call of this thread was dispatched to (mvster_et_last_kernel)."""
    return _lib.load().mvster_et_last_kernel().decode()


def et_normalize(cost: Tensor, wsum: Tensor) -> Tensor:
    B, D, H, W, G = cost.shape
    _chk(cost, "cost")
    _chk(wsum, "wsum", (B, D, H, W))
    _lib.check(_lib.load().mvster_et_normalize_f32(_ptr(cost), _ptr(wsum), B, G, D, H, W, _stream()), "mvster_et_normalize_f32")
    return cost


def et_fuse_bwd(ref: Tensor, srcs: Sequence[Tensor], pose: Tensor, hypo: Tensor, cost: Tensor, wsum: Tensor,
                grad_cost: Tensor, attn_temp: float, need_src: Optional[Sequence[bool]] = None):
    """Gradients of ``et_fuse`` (group correlation, attn_fuse_d) w.r.t. the features: ref [B,H,W,C], srcs V x [B,Hs,Ws,C],
    pose [B,V,12], hypo [B,D,H,W], cost / grad_cost [B,D,H,W,G], wsum [B,D,H,W] (the partial sums, no 1e-8 seed) ->
    (grad_ref [B,H,W,C], [grad_src_v [B,Hs,Ws,C] or None]).  ``need_src[v]`` False skips that view's scatter."""
    B, H, W, Cc = ref.shape
    _chk(ref, "ref")
    V = len(srcs)
    D, G = hypo.shape[1], cost.shape[-1]
    _chk(hypo, "hypo", (B, D, H, W))
    _chk(cost, "cost", (B, D, H, W, G))
    _chk(grad_cost, "grad_cost", (B, D, H, W, G))
    _chk(wsum, "wsum", (B, D, H, W))
    Hs, Ws = srcs[0].shape[1:3]
    for i, s in enumerate(srcs):
        _chk(s, f"src[{i}]", (B, Hs, Ws, Cc))
    need = [True] * V if need_src is None else list(need_src)
    grad_src = [torch.zeros_like(s) if n else None for s, n in zip(srcs, need)]
    grad_ref = torch.zeros_like(ref)
    lib = _lib.load()
    for v0 in range(0, V, MAX_VIEWS):  # the gradient is a sum over views: chunks of MAX_VIEWS add up
        chunk = srcs[v0:v0 + MAX_VIEWS]
        pose_c = _chk(pose[:, v0:v0 + len(chunk)].contiguous(), "pose", (B, len(chunk), 12))
        arr = (C.c_void_p * len(chunk))(*[s.data_ptr() for s in chunk])
        garr = (C.c_void_p * len(chunk))(*[None if g is None else g.data_ptr() for g in grad_src[v0:v0 + len(chunk)]])
        part = grad_ref if v0 == 0 else torch.empty_like(ref)
        _lib.check(lib.mvster_et_fuse_bwd_f32(_ptr(ref), arr, len(chunk), _ptr(pose_c), _ptr(hypo), _ptr(cost), _ptr(wsum),
                                              _ptr(grad_cost), _ptr(part), garr, B, Cc, G, D, H, W, Hs, Ws,
                                              float(attn_temp), _stream()), "mvster_et_fuse_bwd_f32")
        if v0:
            grad_ref += part
    return grad_ref, grad_src


def conv3d_ndhwc(x: Tensor, w: Tensor, bias: Optional[Tensor], kd: int, stride_d: int = 1, stride_hw: int = 1,
                 transposed: bool = False, relu: bool = True, skip: Optional[Tensor] = None) -> Tensor:
    """x [B,D,H,W,Cin], w [kd*9,Cin,Cout] -> y [B,Do,Ho,Wo,Cout]."""
    _chk(x, "x")
    B, Di, Hi, Wi, Cin = x.shape
    taps, cin2, Cout = w.shape
    _chk(w, "w", (kd * 9, Cin, Cout))
    if transposed:
        Do, Ho, Wo = Di, 2 * Hi, 2 * Wi
    else:
        Do, Ho, Wo = (Di - 1) // stride_d + 1, (Hi - 1) // stride_hw + 1, (Wi - 1) // stride_hw + 1
    y = torch.empty((B, Do, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    if bias is not None:
        _chk(bias, "bias", (Cout,))
    if skip is not None:
        _chk(skip, "skip", tuple(y.shape))
    _lib.check(_lib.load().mvster_conv3d_ndhwc_f32(_ptr(x), _ptr(w), _ptr(bias), _ptr(skip), _ptr(y), B, Di, Hi, Wi, Cin, Cout,
                                                   kd, stride_d, stride_hw, int(transposed), int(relu), _stream()),
               "mvster_conv3d_ndhwc_f32")
    return y


def conv3d_tc2(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, kd: int, relu: bool = True,
               skip: Optional[Tensor] = None, npass: int = 3) -> Tensor:
    """Generation-2 tcgen05 conv (tile staged once per plane): w_packed from packing.pack_tc2_weights."""
    _chk(x, "x")
    _chk(w_packed, "w_packed")
    B, D, H, W, Cin = x.shape
    y = torch.empty((B, D, H, W, cout), device=x.device, dtype=torch.float32)
    if bias is not None:
        _chk(bias, "bias", (cout,))
    if skip is not None:
        _chk(skip, "skip", tuple(y.shape))
    want = kd * 9 * Cin * max(cout, 16) * (2 if npass == 3 else 1)
    if w_packed.numel() != want:
        raise ValueError(f"w_packed has {w_packed.numel()} floats, expected {want}")
    _lib.check(_lib.load().mvster_conv3d_tc2_f32(_ptr(x), _ptr(w_packed), _ptr(bias), _ptr(skip), _ptr(y), B, D, H, W, Cin, cout,
                                                 kd, int(relu), npass, _stream()), "mvster_conv3d_tc2_f32")
    return y


def conv_tc3_plan(cin: int, kd: int, k: int, stride: int) -> List[tuple]:
    """(kz, ky, kx, first input channel, ky2, kx2) of every weight slab (= MMA slot) of a generation-3 layer, in the order the
    kernel streams them; (ky2, kx2) is the second tap sharing the MMA of an <= 8-channel layer, else (-1, -1).
    Pure host logic of the library (no GPU needed)."""
    lib = _lib.load()
    n = lib.mvster_conv_tc3_plan(cin, kd, k, stride, None, 0)
    if n < 0:
        raise ValueError(f"conv_tc3: unsupported layer Cin={cin} kd={kd} k={k} stride={stride}")
    buf = (C.c_int * (6 * n))()
    lib.mvster_conv_tc3_plan(cin, kd, k, stride, buf, n)
    return [tuple(buf[6 * i:6 * i + 6]) for i in range(n)]


TC3_FP16X2 = 256  # MVSTER_TC3_FP16X2: the packed weights hold two fp16 terms (packing.pack_tc3_weights(split=2))
TC3_BF16X1 = 512  # MVSTER_TC3_BF16X1: one bf16 term (split=1): the bf16-storage arithmetic
REG2D_SCALE_FLOATS = 288


def _tc3_flags(relu: bool, split: int, round_out: bool = False) -> int:
    if split not in (1, 2, 3):
        raise ValueError(f"split must be 1 (one bf16 term), 2 (two fp16 terms) or 3 (three bf16 terms), got {split}")
    return int(relu) | {1: TC3_BF16X1, 2: TC3_FP16X2, 3: 0}[split] | (2048 if round_out and split == 1 else 0)


def conv_tc3(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, kd: int, k: int, stride: int = 1, relu: bool = True,
             skip: Optional[Tensor] = None, out: Optional[Tensor] = None, split: int = 3, scale: Optional[Tensor] = None) -> Tensor:
    """Generation-3 tcgen05 conv (persistent): x [B,D,H,W,Cin] -> [B,D,Ho,Wo,cout]; w_packed from
    packing.pack_tc3_weights(split=split) (a float32-typed byte blob); split 3 = three bf16 terms per operand, 2 = two fp16,
    1 = one bf16 term (the input is rounded to bf16: bf16-storage arithmetic).  ``scale`` [cout]: per-channel factor applied
    to the accumulator before the bias."""
    _chk(x, "x")
    _chk(w_packed, "w_packed")
    B, D, H, W, Cin = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty((B, D, Ho, Wo, cout), device=x.device, dtype=torch.float32) if out is None else _chk(out, "out", (B, D, Ho, Wo, cout))
    if bias is not None:
        _chk(bias, "bias", (cout,))
    if skip is not None:
        _chk(skip, "skip", tuple(y.shape))
    lib = _lib.load()
    want = lib.mvster_conv_tc3_packed_bytes(Cin, cout, kd, k, stride)
    if want == 0 or w_packed.numel() * 4 != want:
        raise ValueError(f"conv_tc3: w_packed holds {w_packed.numel() * 4} bytes, layer needs {want} (0 = unsupported layer)")
    if scale is not None:
        _chk(scale, "scale", (cout,))
    _lib.check(lib.mvster_conv_tc3_scaled_f32(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(bias), _ptr(skip), _ptr(y), B, D, H, W, Cin, cout,
                                              kd, k, stride, _tc3_flags(relu, split), _stream()), "mvster_conv_tc3_f32")
    return y


def hypo_init_linear(depth_values: Tensor, D: int, H: int, W: int) -> Tensor:
    dv = _chk(depth_values, "depth_values")
    B, n = dv.shape
    out = torch.empty((B, D, H, W), device=dv.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_hypo_init_linear_f32(_ptr(dv), n, _ptr(out), B, D, H, W, _stream()), "mvster_hypo_init_linear_f32")
    return out


def hypo_schedule_linear(depth: Tensor, depth_values: Tensor, ratio: float, D: int, H: int, W: int) -> Tensor:
    """depth [B,H/2,W/2] of the previous stage -> [B,D,H,W] (mvs4net_utils.py:88-99)."""
    B = depth.shape[0]
    _chk(depth, "depth", (B, H // 2, W // 2))
    dv = _chk(depth_values, "depth_values")
    out = torch.empty((B, D, H, W), device=depth.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_hypo_schedule_linear_f32(_ptr(depth), _ptr(dv), dv.shape[1], float(ratio), _ptr(out), B, D, H, W, _stream()),
               "mvster_hypo_schedule_linear_f32")
    return out


def reg3d_layer_table(G: int, down_size: int) -> List[dict]:
    lib = _lib.load()
    out = []
    for i in range(int(lib.mvster_reg3d_num_layers(down_size))):
        info = (C.c_int64 * 8)()
        _lib.check(lib.mvster_reg3d_layer_info(G, down_size, i, info), "mvster_reg3d_layer_info")
        name = "prob" if info[7] == 100 else f"conv{info[7]}"
        out.append(dict(name=name, cin=info[0], cout=info[1], stride=info[2], transposed=bool(info[3]), w_off=info[4], b_off=info[5], taps=info[6]))
    return out


def reg3d(blob: Tensor, cost: Tensor, down_size: int) -> Tensor:
    """cost [B,D,H,W,G] -> logits [B,D,H,W] (the whole reg3d U-Net including its 3x3x3 prob layer)."""
    _chk(cost, "cost")
    _chk(blob, "blob")
    B, D, H, W, G = cost.shape
    lib = _lib.load()
    if blob.numel() != int(lib.mvster_reg3d_blob_floats(G, down_size)):
        raise ValueError(f"blob has {blob.numel()} floats, expected {int(lib.mvster_reg3d_blob_floats(G, down_size))}")
    ws = torch.empty(int(lib.mvster_reg3d_workspace_floats(B, D, H, W)), device=cost.device, dtype=torch.float32)
    out = torch.empty((B, D, H, W), device=cost.device, dtype=torch.float32)
    _lib.check(lib.mvster_reg3d_f32(_ptr(blob), _ptr(cost), _ptr(out), _ptr(ws), B, G, D, H, W, down_size, _stream()), "mvster_reg3d_f32")
    return out


def deconv_tc3(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, rows: int = -1, relu: bool = True,
               skip: Optional[Tensor] = None, out: Optional[Tensor] = None, split: int = 3, scale: Optional[Tensor] = None) -> Tensor:
    """Transposed conv (1,3,3) / stride (1,2,2) on the generation-3 tcgen05 kernel: x [B,D,H,W,Cin] -> [B,D,2H,2W,cout].
    rows = -1 writes every output pixel; rows = 0 / 1 only the output rows of that parity (pass ``out`` to the second call)."""
    _chk(x, "x")
    _chk(w_packed, "w_packed")
    B, D, H, W, Cin = x.shape
    y = torch.empty((B, D, 2 * H, 2 * W, cout), device=x.device, dtype=torch.float32) if out is None else _chk(out, "out", (B, D, 2 * H, 2 * W, cout))
    if bias is not None:
        _chk(bias, "bias", (cout,))
    if skip is not None:
        _chk(skip, "skip", tuple(y.shape))
    lib = _lib.load()
    want = lib.mvster_deconv_tc3_packed_bytes(Cin, cout, rows)
    if want == 0 or w_packed.numel() * 4 != want:
        raise ValueError(f"deconv_tc3: w_packed holds {w_packed.numel() * 4} bytes, layer needs {want} (0 = unsupported layer)")
    if scale is not None:
        _chk(scale, "scale", (cout,))
    _lib.check(lib.mvster_deconv_tc3_scaled_f32(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(bias), _ptr(skip), _ptr(y), B, D, H, W, Cin, cout, rows,
                                                _tc3_flags(relu, split), _stream()), "mvster_deconv_tc3_f32")
    return y


def reg2d_layer_table(G: int) -> List[dict]:
    lib = _lib.load()
    n_layers = 10
    out = []
    for i in range(n_layers):
        info = (C.c_int64 * 8)()
        _lib.check(lib.mvster_reg2d_layer_info(G, i, info), "mvster_reg2d_layer_info")
        out.append(dict(cin=info[0], cout=info[1], kd=info[2], stride=info[3], transposed=bool(info[4]),
                        w_off=info[5], b_off=info[6], taps=info[7]))
    return out


def reg2d_blob_floats(G: int) -> int:
    return int(_lib.load().mvster_reg2d_blob_floats(G))


def reg2d_workspace_floats(B: int, D: int, H: int, W: int) -> int:
    return int(_lib.load().mvster_reg2d_workspace_floats(B, D, H, W))


REG2D_PACKED = 4096  # MVSTER_REG2D_PACKED


def reg2d(blob: Tensor, cost: Tensor, workspace: Optional[Tensor] = None, out: Optional[Tensor] = None,
          tc_blob: Optional[Tensor] = None, npass: int = 3, kernel_gen: int = 2, split: int = 3, packed: bool = False) -> Tensor:
    """cost [B,D,H,W,G] -> feat8 [B,D,H,W,8] (everything of reg2d except the 1x1x1 prob layer).
    With ``tc_blob``: kernel_gen 3 = conv1..conv11 on the persistent tcgen05 kernel (pack_reg2d 'tc3_blob' / 'tc3h_blob' with
    split 3 / 2); kernel_gen 2 = the three 3x3x3 layers on the staged-tile TF32 kernel (pack_tc2_weights slabs; npass 3 = 3xTF32, 1 = TF32)."""
    _chk(cost, "cost")
    _chk(blob, "blob")
    B, D, H, W, G = cost.shape
    need = reg2d_workspace_floats(B, D, H, W)
    if workspace is None:
        workspace = torch.empty(need, device=cost.device, dtype=torch.float32)
    if workspace.numel() < need:
        raise ValueError(f"workspace too small: {workspace.numel()} < {need} floats")
    if blob.numel() != reg2d_blob_floats(G):
        raise ValueError(f"blob has {blob.numel()} floats, expected {reg2d_blob_floats(G)} for G={G}")
    out = torch.empty((B, D, H, W, 8), device=cost.device, dtype=torch.float32) if out is None else _chk(out, "feat8", (B, D, H, W, 8))
    if tc_blob is not None and kernel_gen == 3:  # conv0..conv6 on the persistent 3 x bf16 kernel (packing.pack_reg2d 'tc3_blob')
        _chk(tc_blob, "tc3_blob", (int(_lib.load().mvster_reg2d_tc3_blob_bytes(G)) // 4,))
        _lib.check(_lib.load().mvster_reg2d_tc3_ex_f32(_ptr(blob), _ptr(tc_blob), _ptr(cost), _ptr(out), _ptr(workspace), B, G, D, H, W,
                                                       (TC3_FP16X2 | (REG2D_PACKED if packed else 0)) if split == 2 else 0, _stream()),
                   "mvster_reg2d_tc3_f32")
        return out
    if tc_blob is not None:
        _chk(tc_blob, "tc_blob", (int(_lib.load().mvster_reg2d_tc_blob_floats()),))
        _lib.check(_lib.load().mvster_reg2d_tc_f32(_ptr(blob), _ptr(tc_blob), _ptr(cost), _ptr(out), _ptr(workspace), B, G, D, H, W,
                                                   npass, kernel_gen, _stream()), "mvster_reg2d_tc_f32")
        return out
    _lib.check(_lib.load().mvster_reg2d_f32(_ptr(blob), _ptr(cost), _ptr(out), _ptr(workspace), B, G, D, H, W, _stream()),
               "mvster_reg2d_f32")
    return out


TC3_OUT_PB16 = 1024   # MVSTER_TC3_OUT_PB16
TC3_ROUND_OUT = 2048  # MVSTER_TC3_ROUND_OUT


def to_pb16(x: Tensor) -> Tensor:
    """[P..., H, W, C] fp32 / bf16 (channels last) -> the packed-operand layout of the bf16 tensor-core layers: octet-planar
    bf16 [P..., C/8, H, W, 8] (round to nearest even).  Host-side helper for tests and tools: inside the path the producing
    kernels write this layout themselves."""
    *lead, H, W, Cc = x.shape
    return x.to(torch.bfloat16).reshape(*lead, H, W, Cc // 8, 8).movedim(-2, -4).contiguous()


def from_pb16(x: Tensor) -> Tensor:
    """Inverse of ``to_pb16``: [P..., C/8, H, W, 8] bf16 -> [P..., H, W, C] fp32."""
    *lead, n8, H, W, _ = x.shape
    return x.float().movedim(-4, -2).reshape(*lead, H, W, n8 * 8).contiguous()


def to_ph16(x: Tensor) -> Tensor:
    """[P..., H, W, C] fp32 -> the packed fp16-pair layout of the two-term tensor-core layers: [P..., C/8, 2, H, W, 8] fp16 with
    x == a1 + 2^-11 a2 to 22 bits (a1 = fp16(x), a2 = fp16(2^11 (x - a1)); |x| < 65504).  Host-side helper for tests and tools."""
    from . import packing
    *lead, H, W, Cc = x.shape
    a1, a2 = packing.fp16_split2(x.float())
    t = torch.stack([a1, a2], dim=-1)                                     # [..., H, W, C, 2]
    return t.reshape(*lead, H, W, Cc // 8, 8, 2).permute(*range(len(lead)), -3, -1, -5, -4, -2).contiguous()


def from_ph16(x: Tensor) -> Tensor:
    """Inverse of ``to_ph16``: [P..., C/8, 2, H, W, 8] fp16 -> [P..., H, W, C] fp32 (a1 + 2^-11 a2, exact)."""
    *lead, n8, two, H, W, _ = x.shape
    v = x[..., 0, :, :, :].float() + x[..., 1, :, :, :].float() / 2048.0   # [..., C/8, H, W, 8]
    return v.movedim(-4, -2).reshape(*lead, H, W, n8 * 8).contiguous()


def _packed_shape(x: Tensor):
    """(B, D, Cin, H, W, term planes, dtype) of a packed activation tensor: [B,D,C/8,H,W,8] bf16 or [B,D,C/8,2,H,W,8] fp16."""
    if x.dtype == torch.bfloat16 and x.dim() == 6:
        B, D, n8, H, W, _ = x.shape
        return B, D, 8 * n8, H, W, 1
    if x.dtype == torch.float16 and x.dim() == 7 and x.shape[3] == 2:
        B, D, n8, _, H, W, _ = x.shape
        return B, D, 8 * n8, H, W, 2
    raise TypeError(f"packed activations must be [B,D,C/8,H,W,8] bfloat16 or [B,D,C/8,2,H,W,8] float16, got {tuple(x.shape)} {x.dtype}")


def _packed_out(B, D, cout, Ho, Wo, NT):
    return ((B, D, cout // 8, Ho, Wo, 8), torch.bfloat16) if NT == 1 else ((B, D, cout // 8, 2, Ho, Wo, 8), torch.float16)


def conv_tc3_pb16(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, kd: int, k: int, stride: int = 1, relu: bool = True,
                  skip: Optional[Tensor] = None, scale: Optional[Tensor] = None, out_pb16: bool = True) -> Tensor:
    """Packed-operand tcgen05 conv (mvster_conv_tc3_pb16): x [B,D,Cin/8,H,W,8] bf16 (one bf16 term; w_packed split = 1) or
    [B,D,Cin/8,2,H,W,8] fp16 (the fp16 pair; split = 2) -> the same packing of the output (out_pb16) or [B,D,Ho,Wo,cout] fp32;
    skip in the packed layout of the output."""
    _chk(x, "x", dtype=x.dtype)
    _chk(w_packed, "w_packed")
    B, D, Cin, H, W, NT = _packed_shape(x)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    pshape, pdt = _packed_out(B, D, cout, Ho, Wo, NT)
    if out_pb16:
        y = torch.empty(pshape, device=x.device, dtype=pdt)
    else:
        y = torch.empty((B, D, Ho, Wo, cout), device=x.device, dtype=torch.float32)
    if skip is not None:
        _chk(skip, "skip", pshape, pdt)
    for t, n in ((bias, "bias"), (scale, "scale")):
        if t is not None:
            _chk(t, n, (cout,))
    lib = _lib.load()
    want = lib.mvster_conv_tc3_packed_bytes(Cin, cout, kd, k, stride)
    if want == 0 or w_packed.numel() * 4 != want:
        raise ValueError(f"conv_tc3_pb16: w_packed holds {w_packed.numel() * 4} bytes, layer needs {want} (0 = unsupported layer)")
    _lib.check(lib.mvster_conv_tc3_pb16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(bias), _ptr(skip), _ptr(y), B, D, H, W, Cin, cout,
                                        kd, k, stride, int(relu) | (TC3_OUT_PB16 if out_pb16 else 0) | (TC3_FP16X2 if NT == 2 else 0), _stream()),
               "mvster_conv_tc3_pb16")
    return y


def deconv_tc3_pb16(x: Tensor, w_packed: Tensor, bias: Optional[Tensor], cout: int, rows: int = -1, relu: bool = True,
                    skip: Optional[Tensor] = None, scale: Optional[Tensor] = None, out_pb16: bool = True, out: Optional[Tensor] = None) -> Tensor:
    """Packed-operand transposed conv (mvster_deconv_tc3_pb16): x packed as in ``conv_tc3_pb16`` -> the same packing at
    (2H, 2W) or [B,D,2H,2W,cout] fp32; rows as in ``deconv_tc3`` (pass ``out`` to the second call)."""
    _chk(x, "x", dtype=x.dtype)
    _chk(w_packed, "w_packed")
    B, D, Cin, H, W, NT = _packed_shape(x)
    pshape, pdt = _packed_out(B, D, cout, 2 * H, 2 * W, NT)
    shape, dt = (pshape, pdt) if out_pb16 else ((B, D, 2 * H, 2 * W, cout), torch.float32)
    y = torch.empty(shape, device=x.device, dtype=dt) if out is None else _chk(out, "out", shape, dt)
    if skip is not None:
        _chk(skip, "skip", pshape, pdt)
    for t, n in ((bias, "bias"), (scale, "scale")):
        if t is not None:
            _chk(t, n, (cout,))
    lib = _lib.load()
    want = lib.mvster_deconv_tc3_packed_bytes(Cin, cout, rows)
    if want == 0 or w_packed.numel() * 4 != want:
        raise ValueError(f"deconv_tc3_pb16: w_packed holds {w_packed.numel() * 4} bytes, layer needs {want} (0 = unsupported layer)")
    _lib.check(lib.mvster_deconv_tc3_pb16(_ptr(x), _ptr(w_packed), _ptr(scale), _ptr(bias), _ptr(skip), _ptr(y), B, D, H, W, Cin, cout, rows,
                                          int(relu) | (TC3_OUT_PB16 if out_pb16 else 0) | (TC3_FP16X2 if NT == 2 else 0), _stream()),
               "mvster_deconv_tc3_pb16")
    return y


def reg2d_bf16(blob_q: Tensor, tc3_blob: Tensor, scales: Tensor, cost: Tensor, workspace: Optional[Tensor] = None,
               out: Optional[Tensor] = None, packed: bool = True) -> Tensor:
    """bf16-storage reg2d (mvster_reg2d_bf16): cost [B,D,H,W,G] torch.bfloat16 -> feat8 [B,D,H,W,8] fp32; the three weight
    tensors come from packing.pack_reg2d_bf16.  ``packed``: activations between the layers as packed bf16 operands
    (mvster_conv_tc3_pb16) instead of bf16-rounded values in fp32 containers - same arithmetic."""
    _chk(cost, "cost", dtype=torch.bfloat16)
    B, D, H, W, G = cost.shape
    _chk(blob_q, "blob_q", (reg2d_blob_floats(G),))
    _chk(tc3_blob, "tc3_blob", (int(_lib.load().mvster_reg2d_tc3_blob_bytes(G)) // 4,))
    _chk(scales, "scales", (REG2D_SCALE_FLOATS,))
    need = reg2d_workspace_floats(B, D, H, W)
    if workspace is None:
        workspace = torch.empty(need, device=cost.device, dtype=torch.float32)
    if workspace.numel() < need:
        raise ValueError(f"workspace too small: {workspace.numel()} < {need} floats")
    out = torch.empty((B, D, H, W, 8), device=cost.device, dtype=torch.float32) if out is None else _chk(out, "feat8", (B, D, H, W, 8))
    _lib.check(_lib.load().mvster_reg2d_bf16(_ptr(blob_q), _ptr(tc3_blob), _ptr(scales), _ptr(cost), _ptr(out), _ptr(workspace),
                                             B, G, D, H, W, 1 if packed else 0, _stream()), "mvster_reg2d_bf16")
    return out


def head(hypo: Tensor, split_itv: float, logits: Optional[Tensor] = None, feat8: Optional[Tensor] = None,
         prob_w: Optional[Tensor] = None, prob_b: Optional[Tensor] = None, inverse: bool = True,
         want_soft: bool = False, bf16_input: bool = False) -> dict:
    _chk(hypo, "hypo")
    B, D, H, W = hypo.shape
    if logits is not None:
        _chk(logits, "logits", (B, D, H, W))
    else:
        _chk(feat8, "feat8", (B, D, H, W, 8))
        _chk(prob_w, "prob_w", (8,))
        _chk(prob_b, "prob_b", (1,))
    dev = hypo.device
    new = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    out = {"attn_weight": new(B, D, H, W), "depth": new(B, H, W), "conf_low": new(B, H, W)}
    if inverse:
        out["inverse_min_depth"] = new(B, H, W)
        out["inverse_max_depth"] = new(B, H, W)
    if want_soft:
        out["soft_depth"] = new(B, H, W)
    _lib.check(_lib.load().mvster_head_ex_f32(_ptr(logits), _ptr(feat8), _ptr(prob_w), _ptr(prob_b), _ptr(hypo),
                                              _ptr(out["attn_weight"]), _ptr(out["depth"]), _ptr(out["conf_low"]),
                                              _ptr(out.get("inverse_min_depth")), _ptr(out.get("inverse_max_depth")),
                                              _ptr(out.get("soft_depth")), B, D, H, W, float(split_itv), 1 if bf16_input else 0, _stream()),
               "mvster_head_f32")
    return out


def upsample_bilinear(x: Tensor, factor: int) -> Tensor:
    _chk(x, "x")
    B, H, W = x.shape
    if factor == 1:
        return x.clone()
    out = torch.empty((B, H * factor, W * factor), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_upsample_bilinear_f32(_ptr(x), _ptr(out), B, H, W, factor, _stream()),
               "mvster_upsample_bilinear_f32")
    return out


def sinkhorn(gt_depth: Tensor, hypo: Tensor, attn: Tensor, mask: Tensor, iters: int, eps: float, continuous: bool, want_grad: bool = True):
    """mvster_sinkhorn_f32: per-pixel transport cost [B,H,W] (0 outside the mask) and, with ``want_grad``, its gradient w.r.t.
    ``attn`` [B,D,H,W].  mask: torch.bool [B,H,W]."""
    B, D, H, W = attn.shape
    _chk(attn, "attn")
    _chk(hypo, "hypo", (B, D, H, W))
    _chk(gt_depth, "gt_depth", (B, H, W))
    _chk(mask, "mask", (B, H, W), torch.bool)
    loss_px = torch.empty((B, H, W), device=attn.device, dtype=torch.float32)
    grad = torch.empty((B, D, H, W), device=attn.device, dtype=torch.float32) if want_grad else None
    _lib.check(_lib.load().mvster_sinkhorn_f32(_ptr(gt_depth), _ptr(hypo), _ptr(attn), _ptr(mask), _ptr(loss_px), _ptr(grad), B, D, H, W,
                                               int(iters), float(eps), int(bool(continuous)), _stream()), "mvster_sinkhorn_f32")
    return loss_px, grad

