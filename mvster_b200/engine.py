"""Inference engine: drives the 4-stage cascade (MVS4Net.py:78-105) through the C ABI.

One engine per CUDA device.  It owns the folded/packed regulariser weights for that device
and nothing else; feature extraction (FPN4, outside the named hot path) runs through the
module's own ``feature`` sub-module in channels-last memory format so its outputs are
already the NHWC tensors the kernels consume.

Configuration coverage: reg2d and reg3d; group correlation and per-channel squared difference
(``group_cor=False``); ``attn_fuse_d`` on/off; inverse and linear depth sampling.
"""
from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, capi, fpn_engine, packing

Tensor = torch.Tensor
# One CUDA-graph capture at a time per process: replicas driven by different host threads (nn.DataParallel) may replay
# concurrently, but two captures running side by side trip over each other in the caching allocator.
_CAPTURE_LOCK = threading.Lock()
REG3D_DOWN = (3, 3, 2, 2)  # MVS4Net.py:48


class StagePlan:
    """Static per-stage configuration lifted from the module."""

    def __init__(self, k: int, net):
        self.k = k
        self.D = net.stage_splits[k]
        self.C = net.feature.out_channels[k]
        self.group_cor = bool(net.group_cor)
        self.G = net.group_cor_dim[k] if self.group_cor else self.C  # channels of the cost volume
        self.split_itv = float(net.depth_interals_ratio[k])
        self.up = 2 ** (3 - k)  # confidence up-sampling factor, mvs4net_utils.py:1077
        self.down = REG3D_DOWN[k] if net.reg_net == "reg3d" else 0


def check_supported(net) -> None:
    bad = []
    if net.num_stage != 4:
        bad.append(f"num_stage={net.num_stage}")
    for k in range(min(net.num_stage, 4)):
        if net.stage_splits[k] not in (4, 8):
            bad.append(f"stage_splits[{k}]={net.stage_splits[k]} (4 or 8)")
        if net.group_cor and net.group_cor_dim[k] not in (4, 8):
            bad.append(f"group_cor_dim[{k}]={net.group_cor_dim[k]} (4 or 8)")
    # widths the kernels and the weight packers are built for (the reference itself breaks for reg_channel != 8:
    # reg2d.prob hard-codes 8 input channels, mvs4net_utils.py:900)
    if getattr(net.feature, "base_channels", 8) != 8 and getattr(net, "fpn_backend", "native") == "native":
        bad.append(f"fpn_base_channel={net.feature.base_channels} with the native feature pyramid (8; set fpn_backend='torch')")
    if list(net.feature.out_channels[:4]) != [64, 32, 16, 8]:
        bad.append(f"feature.out_channels={list(net.feature.out_channels)} ([64, 32, 16, 8])")
    reg0 = net.reg[0] if len(net.reg) else None
    base = getattr(getattr(getattr(reg0, "conv0", None), "conv", None), "out_channels", 8)
    if base != 8:
        bad.append(f"reg_channel={base} (8)")
    if bad:
        raise NotImplementedError("the CUDA inference path does not cover: " + ", ".join(bad))


BF16_ET_CLASSES = ((64, 8, 8), (32, 8, 8), (16, 4, 4), (8, 4, 4))  # (C, G, D) the bf16 warp / ET kernels are instantiated for


def bf16_storage(net) -> bool:
    st = getattr(net, "storage", "fp32")
    if st not in ("fp32", "bf16"):
        raise ValueError(f"storage must be 'fp32' or 'bf16', got {st!r}")
    return st == "bf16"


def check_bf16_supported(net, shard) -> None:
    bad = []
    if net.reg_net != "reg2d":
        bad.append(f"reg_net={net.reg_net!r} (reg2d)")
    if not net.group_cor:
        bad.append("group_cor=False")
    if not net.stagenet.attn_fuse_d:
        bad.append("attn_fuse_d=False")
    if getattr(net, "fpn_backend", "native") != "native":
        bad.append("fpn_backend != 'native'")
    if shard is not None:
        bad.append("view sharding (partial sums are fp32)")
    for k in range(min(net.num_stage, 4)):
        cls = (net.feature.out_channels[k], net.group_cor_dim[k], net.stage_splits[k])
        if net.group_cor and cls not in BF16_ET_CLASSES:
            bad.append(f"stage {k + 1} (C, G, D) = {cls}")
    if bad:
        raise NotImplementedError("storage='bf16' does not cover: " + ", ".join(bad))


def check_inputs(net, imgs: Sequence[Tensor], proj_matrices: Dict[str, Tensor], depth_values: Tensor) -> None:
    """Shape contract of ``MVS4net.forward`` (MVS4Net.py:60-77), checked before anything is launched; the reference
    fails on the same inputs with ATen shape errors deep inside the U-Nets."""
    if len(imgs) < 2:
        raise ValueError(f"MVS4net.forward needs the reference view and at least one source view, got {len(imgs)} image tensor(s)")
    shape = tuple(imgs[0].shape)
    if len(shape) != 4 or shape[1] != 3 or any(tuple(i.shape) != shape for i in imgs):
        raise ValueError(f"imgs must be Nv tensors of one shape [B,3,H,W], got {[tuple(i.shape) for i in imgs]}")
    B, _, H, W = shape
    if H % 64 or W % 64 or H == 0 or W == 0:
        raise ValueError(f"H and W must be multiples of 64 (four pyramid levels, three stride-2 levels in the regulariser), got {H}x{W}")
    for k in range(net.num_stage):
        key = f"stage{k + 1}"
        if key not in proj_matrices or tuple(proj_matrices[key].shape) != (B, len(imgs), 2, 4, 4):
            got = tuple(proj_matrices[key].shape) if key in proj_matrices else None
            raise ValueError(f"proj_matrices['{key}'] must be [B={B}, Nv={len(imgs)}, 2, 4, 4], got {got}")
    if depth_values.dim() != 2 or depth_values.shape[0] != B or depth_values.shape[1] < 2:
        raise ValueError(f"depth_values must be [B={B}, >= 2], got {tuple(depth_values.shape)}")


def module_state(net) -> Dict[str, Tensor]:
    """``state_dict()`` that also works on an nn.DataParallel replica: `replicate` keeps a replica's parameters as plain tensor
    attributes (listed in ``_former_parameters``), so its own ``state_dict()`` only holds the buffers."""
    out = {k: v.detach() for k, v in net.state_dict().items()}
    for prefix, mod in net.named_modules():
        former = getattr(mod, "_former_parameters", None)
        if former:
            for k, v in former.items():
                out[(prefix + "." if prefix else "") + k] = v.detach()
    return out


def _dup(v, memo: Dict[int, Tensor], srcs: List[Tensor], dsts: List[Tensor]):
    if isinstance(v, dict):
        return {k: _dup(x, memo, srcs, dsts) for k, x in v.items()}
    if not torch.is_tensor(v):
        return v
    got = memo.get(id(v))
    if got is None:
        got = memo[id(v)] = torch.empty_like(v)  # preserve_format: the permuted mono_feat view keeps its strides
        srcs.append(v)
        dsts.append(got)
    return got


def _copy_outputs(out: Dict) -> Dict:
    """Fresh tensors for every distinct tensor of an output dict (one multi-tensor copy); entries that alias each other in
    ``out`` (the top-level keys and ``stage4``'s, MVS4Net.py:104-105) alias each other in the result.  (A module-level helper, not
    a recursive closure: a closure that refers to itself is cyclic garbage that keeps its lists - and with them every copied
    tensor, ~45 MB per call at cfg2 - alive until the cycle collector runs; with the collector off the caching allocator then
    falls back to cudaMalloc after a dozen calls and each step stalls for milliseconds - measured at 8 ranks, profiles/r02_multi_gpu.md.)"""
    srcs: List[Tensor] = []
    dsts: List[Tensor] = []
    res = _dup(out, {}, srcs, dsts)
    if srcs:
        torch._foreach_copy_(dsts, srcs)
    return res


class InferenceEngine:
    def __init__(self, device: torch.device):
        self.device = device
        self.weights_version = -1
        self.stage_weights: List[Dict[str, Tensor]] = []
        self.plans: List[StagePlan] = []
        self._graphs: Dict = {}
        self._side = None  # second stream for the early cascade stages (_forward_overlapped)
        self.fp16_safe = {"fpn": True, "reg": True}  # folded weights inside the two-fp16-term range (refresh_weights)
        self._warned_fp16 = False
        # activation range of the two-fp16-term arithmetic: checked on the device during the first forward after the weights
        # changed (and on every eager forward with MVSTER_RANGE_CHECK=always); a hit switches this engine to three bf16 terms
        self._range_checked = False
        self._overflow_flag: Optional[Tensor] = None
        self._reg_state: Dict[str, Tensor] = {}    # CPU view of the regulariser's state: packed for storage='bf16' on first use
        self._bf16_weights: Dict[int, Dict[str, Tensor]] = {}

    def _precision(self, net, which: str) -> str:
        """Arithmetic of the convolutions of ``which`` ("fpn" | "reg"): the module's setting, except that "2xfp16" becomes "3xbf16"
        (same kernel, full fp32 range) for a model whose folded weights do not fit the fp16 terms."""
        prec = getattr(net, which + "_precision", "fp32")
        if prec == "2xfp16" and not self.fp16_safe.get(which, True):
            if not self._warned_fp16:
                import warnings
                warnings.warn(f"mvster_b200: folded {which} weights exceed the range of the two-fp16-term arithmetic "
                              f"(|w| > {packing.FP16_WEIGHT_LIMIT:g} or non-finite); using three bf16 terms instead")
                self._warned_fp16 = True
            return "3xbf16"
        return prec

    # ------------------------------------------------------------------ weights
    def refresh_weights(self, net) -> None:
        check_supported(net)
        self._graphs.clear()  # captured graphs reference the previous packed weights
        self._range_checked = False
        state = module_state(net)
        sd = {k: v for k, v in state.items() if k.startswith("reg.")}
        self._reg_state, self._bf16_weights = sd, {}
        self.plans = [StagePlan(k, net) for k in range(net.num_stage)]
        self.stage_weights = []
        fpn_sd = {k: v for k, v in state.items() if k.startswith("feature.")}
        # Group-interleaved feature channels (MVSTER_ET_INTERLEAVED): the native pyramid writes stage k's channels in the order the
        # window kernels want (a 64-bit pair = the same channel index of two neighbouring groups), by permuting the output
        # channels of out{k} when the weights are packed.  Only for the shipped (C, G, D) classes those kernels are built for.
        self.interleave = [False] * len(self.plans)
        if net.group_cor and getattr(net, "fpn_backend", "torch") == "native" and os.environ.get("MVSTER_ET_INTERLEAVE", "1") == "1":
            self.interleave = [(p.C, p.G, p.D) in ((8, 4, 4), (16, 4, 4), (32, 8, 8)) for p in self.plans]
        self.perms = [capi.interleave_perm(p.C, p.G) if il else None for p, il in zip(self.plans, self.interleave)]
        fpn_packed = fpn_engine.pack_fpn(fpn_sd, out_perm={k + 1: perm for k, perm in enumerate(self.perms)})
        self.fp16_safe = {"fpn": packing.fp16_range_ok(v for k, v in fpn_packed.items() if k.endswith(".w")), "reg": True}
        self.fpn_weights = {k: v.to(self.device) for k, v in fpn_packed.items()}
        with torch.cuda.device(self.device):
            for p in self.plans:
                if net.reg_net == "reg2d":
                    packed = packing.pack_reg2d(sd, f"reg.{p.k}", capi.reg2d_layer_table(p.G))
                else:
                    packed = {"blob": packing.pack_reg3d(sd, f"reg.{p.k}", capi.reg3d_layer_table(p.G, p.down))}
                self.fp16_safe["reg"] = self.fp16_safe["reg"] and packing.fp16_range_ok([packed["blob"]])
                self.stage_weights.append({k: v.to(self.device) for k, v in packed.items()})

    # ------------------------------------------------------------------ full forward
    def forward(self, net, imgs: Sequence[Tensor], proj_matrices: Dict[str, Tensor], depth_values: Tensor,
                shard=None) -> Dict:
        """``shard`` (sharding.ViewShard) restricts this rank to the reference view plus its own
        slice of the source views: features are extracted only for those."""
        check_inputs(net, imgs, proj_matrices, depth_values)
        B = imgs[0].shape[0]
        own = list(range(len(imgs))) if shard is None else [0] + shard.views
        if bf16_storage(net):
            check_bf16_supported(net, shard)
        self._native_feats = getattr(net, "fpn_backend", "torch") == "native"  # features in the pyramid's (interleaved) channel order
        try:
            if self._needs_range_check(net):
                return self._range_checked_forward(net, imgs, proj_matrices, depth_values, shard, B, own)
            return self._forward(net, imgs, proj_matrices, depth_values, shard, B, own)
        finally:
            self._native_feats = False

    def _needs_range_check(self, net) -> bool:
        if self.device.type != "cuda" or torch.cuda.is_current_stream_capturing():
            return False
        if "2xfp16" not in (self._precision(net, "fpn"), self._precision(net, "reg")):
            return False
        return (not self._range_checked) or os.environ.get("MVSTER_RANGE_CHECK", "once") == "always"

    def _range_checked_forward(self, net, imgs, proj_matrices, depth_values, shard, B, own) -> Dict:
        """One forward with the converters of every two-fp16-term convolution reporting inputs that do not fit the fp16 terms
        (|x| >= 65504 or non-finite: `cvt.rn.satfinite` would clamp them silently).  If any does, the engine switches to three
        bf16 terms - same kernel, full fp32 range - warns once and repeats the forward.  Costs one device-to-host read."""
        lib = _lib.load()
        with torch.cuda.device(self.device):
            if self._overflow_flag is None:
                self._overflow_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._overflow_flag.zero_()
            lib.mvster_tc3_set_overflow_flag(self._overflow_flag.data_ptr())
            try:
                out = self._forward(net, imgs, proj_matrices, depth_values, shard, B, own)
            finally:
                lib.mvster_tc3_set_overflow_flag(None)
            hit = int(self._overflow_flag.item()) != 0
        self._range_checked = True
        if hit:
            import warnings
            warnings.warn("mvster_b200: an activation exceeds the range of the two-fp16-term convolution arithmetic (|x| >= 65504 or "
                          "non-finite); this engine now uses three bf16 terms (full fp32 range) instead")
            self.fp16_safe = {"fpn": False, "reg": False}
            self._warned_fp16 = True
            self._graphs.clear()
            out = self._forward(net, imgs, proj_matrices, depth_values, shard, B, own)
        return out

    def _forward(self, net, imgs, proj_matrices, depth_values, shard, B, own) -> Dict:
        with torch.cuda.device(self.device):
            if getattr(net, "fpn_backend", "torch") == "native":
                # FPN4 inside libmvster_b200 (fpn_engine.py): NCHW images in, NHWC features out
                x = torch.cat([imgs[v] for v in own], 0).to(dtype=torch.float32).contiguous()
                prec = self._precision(net, "fpn")
                npass = {"fp32": 0, "3xtf32": 3, "tf32": 1, "3xbf16": 3, "2xfp16": 2}[prec]
                gen = 3 if prec in ("3xbf16", "2xfp16") else 2
                if net.num_stage == 4 and getattr(net, "overlap_stages", True) and (shard is None or shard.count > 0):
                    return self._forward_overlapped(net, x, B, len(own), proj_matrices, depth_values, npass, gen, shard)
                pyramid = fpn_engine.run_fpn(self.fpn_weights, x, npass, gen=gen)
                nhwc = [pyramid[f"stage{k + 1}"] for k in range(net.num_stage)]
                if bf16_storage(net):  # the pyramid's outputs as they are stored: bf16
                    nhwc = [capi.cast_bf16(t) for t in nhwc]
            else:
                x = torch.cat([imgs[v] for v in own], 0).contiguous(memory_format=torch.channels_last)
                pyramid = net.feature(x)  # {stage: [len(own)*B, C, h, w]}, channels-last strides
                nhwc = [capi.to_nhwc(pyramid[f"stage{k + 1}"]) for k in range(net.num_stage)]
            feats = [[f[i * B:(i + 1) * B] for i in range(len(own))] for f in nhwc]
            return self.run_cascade(net, feats, proj_matrices, depth_values, shard=shard)

    # ------------------------------------------------------------------ CUDA-graph replay (the default inference path)
    def forward_graphed(self, net, imgs: Sequence[Tensor], proj_matrices: Dict[str, Tensor], depth_values: Tensor,
                        shard=None) -> Dict:
        """Capture the whole forward (feature pyramid + 4-stage cascade, ~80 launches, plus the per-stage NCCL all-reduce of a
        view-sharded run) into one CUDA graph per input signature and replay it: removes the per-launch host overhead and the
        inter-kernel gaps.  Inputs are copied into static buffers.  The graph's outputs are static buffers too; by default the
        caller gets COPIES (one fused multi-tensor copy after the replay, aliasing between the top-level and the ``stage4``
        entries preserved), so results stay valid across calls like the reference's; ``net.graph_static_outputs = True`` returns
        the static buffers themselves (valid until the next call with the same signature - what a serving loop that consumes
        each frame before requesting the next one wants).  Capture uses the thread-local error mode: nn.DataParallel replicas
        (test_mvs4.py:196) capture and replay from their own host threads."""
        key = (tuple(tuple(t.shape) for t in imgs), tuple(sorted((k, tuple(v.shape)) for k, v in proj_matrices.items())),
               tuple(depth_values.shape), self.weights_version, self._precision(net, "reg"),
               getattr(net, "fpn_backend", "torch"), self._precision(net, "fpn"),
               getattr(net, "storage", "fp32"), os.environ.get("MVSTER_REG_PACKED", "1"), os.environ.get("MVSTER_BF16_PACKED", "1"),
               getattr(net, "overlap_stages", True), os.environ.get("MVSTER_SIDE_SMS", "0"), os.environ.get("MVSTER_MAIN_RESERVE", "48"),
               None if shard is None else (shard.first_view, shard.count, shard.parts, id(shard.group)),
               int(getattr(net, "graph_slot", 0)))
        entry = self._graphs.get(key)
        if entry is None:
            with _CAPTURE_LOCK, torch.cuda.device(self.device):
                s_imgs = [t.detach().to(self.device, torch.float32).clone() for t in imgs]
                s_proj = {k: v.detach().to(self.device, torch.float32).clone() for k, v in proj_matrices.items()}
                s_dv = depth_values.detach().to(self.device, torch.float32).clone()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):  # warm-up outside capture: cuDNN autotune, lazy module loads, func attributes
                    for _ in range(2):
                        self.forward(net, s_imgs, s_proj, s_dv, shard=shard)
                torch.cuda.current_stream(self.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                # an explicit capture stream on THIS device: torch.cuda.graph's default one is created once, on whichever device was
                # current then, and capturing it from another device's replica fails ("operation not permitted when stream is capturing")
                with torch.cuda.graph(graph, stream=torch.cuda.Stream(device=self.device), capture_error_mode="thread_local"):
                    out = self.forward(net, s_imgs, s_proj, s_dv, shard=shard)
            entry = self._graphs[key] = (graph, s_imgs, s_proj, s_dv, out)
            if len(self._graphs) > 4:  # keep the cache small: each graph pins its own workspace
                self._graphs.pop(next(iter(self._graphs)))
        graph, s_imgs, s_proj, s_dv, out = entry
        with torch.cuda.device(self.device):
            for dst, src in zip(s_imgs, imgs):
                dst.copy_(src, non_blocking=True)
            for k, dst in s_proj.items():
                dst.copy_(proj_matrices[k], non_blocking=True)
            s_dv.copy_(depth_values, non_blocking=True)
            graph.replay()
            if getattr(net, "graph_static_outputs", False):
                return out
            return _copy_outputs(out)

    # ------------------------------------------------------------------ the hot path
    def _aggregate(self, p: StagePlan, ref: Tensor, srcs: List[Tensor], proj: Tensor, hypo: Tensor, temp: float,
                   fuse_d: bool, shard, pose: Optional[Tensor] = None) -> Tensor:
        if ref.dtype == torch.bfloat16:  # storage='bf16': bf16 features in, bf16 cost volume out
            return capi.et_fuse_bf16(ref, srcs, capi.pose(proj) if pose is None else pose, hypo, p.G, temp, interleaved=self._interleaved(p.k, ref))
        kw = dict(group_cor=p.group_cor, fuse_d=fuse_d, interleaved=self._interleaved(p.k, ref))
        if shard is None:
            return capi.et_fuse(ref, srcs, capi.pose(proj) if pose is None else pose, hypo, p.G, temp, **kw)
        from . import sharding
        B, H, W, _ = ref.shape

        def partial(acc, wsum, pose=pose):
            if pose is None:
                pose = capi.pose(proj, first_view=shard.first_view, n_views=shard.count)
            capi.et_fuse(ref, srcs, pose, hypo, p.G, temp, cost=acc, wsum=wsum, partial=True, **kw)

        return sharding.sharded_aggregate(partial, capi.et_normalize, (B, p.D, H, W, p.G), shard, self.device)

    def _interleaved(self, k: int, ref: Tensor) -> bool:
        """True when the stage-k features this engine was handed are in the group-interleaved layout (they are whenever they
        come from the engine's own native pyramid, see refresh_weights; ``run_cascade`` callers pass natural-order features)."""
        return bool(getattr(self, "_native_feats", False) and k < len(self.interleave) and self.interleave[k])

    def _regularise(self, net, p: StagePlan, wts: Dict[str, Tensor], cost: Tensor, hypo: Tensor) -> Dict[str, Tensor]:
        inverse = bool(net.inverse_depth)
        if cost.dtype == torch.bfloat16:  # storage='bf16': one-term bf16 operands, BatchNorm scale in fp32 (mvster_reg2d_bf16)
            q = self._bf16_stage_weights(p)
            # MVSTER_BF16_PACKED=0: activations as bf16-rounded values in fp32 containers (converter path) instead of packed operands
            feat8 = capi.reg2d_bf16(q["blob_q"], q["tc3_blob"], q["scales"], cost, packed=os.environ.get("MVSTER_BF16_PACKED", "1") == "1")
            return capi.head(hypo, p.split_itv, feat8=feat8, prob_w=q["prob_w"], prob_b=q["prob_b"], inverse=inverse, bf16_input=True)
        if net.reg_net == "reg3d":
            logits = capi.reg3d(wts["blob"], cost, p.down)
            return capi.head(hypo, p.split_itv, logits=logits, inverse=inverse)
        prec = self._precision(net, "reg")
        if prec == "fp32":      # exact fp32 FMA on the CUDA cores for every layer
            feat8 = capi.reg2d(wts["blob"], cost)
        elif prec == "3xbf16":  # conv0..conv6 on the persistent tcgen05 kernel, three bf16 terms per operand (fp32-faithful)
            feat8 = capi.reg2d(wts["blob"], cost, tc_blob=wts["tc3_blob"], kernel_gen=3)
        elif prec == "2xfp16":  # same kernel, two fp16 terms per operand (22-bit operands, 2/3 of the MMAs; |x| < 65504)
            # MVSTER_REG_PACKED=0: fp32 activations between the layers, split into their fp16 terms by every consuming layer's
            # converter warps (the round-1 form) instead of once by the producing epilogue
            feat8 = capi.reg2d(wts["blob"], cost, tc_blob=wts["tc3h_blob"], kernel_gen=3, split=2,
                               packed=os.environ.get("MVSTER_REG_PACKED", "1") == "1")
        else:                   # 3x3x3 layers on tcgen05: "3xtf32" (fp32-faithful) or "tf32"
            feat8 = capi.reg2d(wts["blob"], cost, tc_blob=wts["tc2_blob"], npass=3 if prec == "3xtf32" else 1, kernel_gen=2)
        return capi.head(hypo, p.split_itv, feat8=feat8, prob_w=wts["prob_w"], prob_b=wts["prob_b"], inverse=inverse)

    def _bf16_stage_weights(self, p: StagePlan) -> Dict[str, Tensor]:
        got = self._bf16_weights.get(p.k)
        if got is None:
            packed = packing.pack_reg2d_bf16(self._reg_state, f"reg.{p.k}", capi.reg2d_layer_table(p.G))
            got = self._bf16_weights[p.k] = {k: v.to(self.device) for k, v in packed.items()}
        return got

    def _run_stage(self, net, p: StagePlan, wts: Dict[str, Tensor], feats_k: List[Tensor], proj_matrices: Dict[str, Tensor],
                   dv: Tensor, prev: Optional[Dict], temp: float, shard=None, pose: Optional[Tensor] = None,
                   deferred: Optional[List] = None, mono_feat: Optional[Tensor] = None) -> Dict:
        """One cascade stage (MVS4Net.py:78-105 loop body + stagenet.forward) on the current stream.  ``pose``: relative poses
        computed earlier (they only depend on the inputs); ``deferred``: if given, the confidence up-sampling - which no later
        stage reads - is appended to it as a closure instead of being launched here, to keep it off the stage-to-stage chain."""
        fuse_d = bool(net.stagenet.attn_fuse_d)
        inverse = bool(net.inverse_depth)
        ref, srcs = feats_k[0], feats_k[1:]
        B, H, W, C = ref.shape
        proj = proj_matrices[f"stage{p.k + 1}"].to(device=self.device, dtype=torch.float32).contiguous()
        if p.k == 0:
            hypo = capi.hypo_init_inverse(dv, p.D, H, W) if inverse else capi.hypo_init_linear(dv, p.D, H, W)
        elif inverse:
            hypo = capi.hypo_schedule_inverse(prev["inverse_min_depth"], prev["inverse_max_depth"], p.D, H, W)
        else:
            hypo = capi.hypo_schedule_linear(prev["depth"], dv, p.split_itv, p.D, H, W)
        cost = self._aggregate(p, ref, srcs, proj, hypo, temp, fuse_d, shard, pose)
        h = self._regularise(net, p, wts, cost, hypo)
        out = {"depth": h["depth"], "hypo_depth": hypo, "attn_weight": h["attn_weight"]}
        if deferred is None:
            out["photometric_confidence"] = capi.upsample_bilinear(h["conf_low"], p.up)
        else:
            deferred.append(lambda o=out, c=h["conf_low"], f=p.up: o.__setitem__("photometric_confidence", capi.upsample_bilinear(c, f)))
        if inverse:
            out["inverse_min_depth"] = h["inverse_min_depth"]
            out["inverse_max_depth"] = h["inverse_max_depth"]
        if net.mono:  # [B,C,H,W] view of the reference features, as mvs4net_utils.py:1092
            out["mono_feat"] = mono_feat if mono_feat is not None else self._mono_feat(p.k, ref)
        return out

    def _mono_feat(self, k: int, ref: Tensor) -> Tensor:
        """The reference view's stage-k features as [B,C,H,W] in the NATURAL channel order (an output only).  Where the engine keeps
        them group-interleaved this is one gather: natural[..., c] = stored[..., inverse[c]]."""
        if ref.dtype != torch.float32:  # storage='bf16': the stored (rounded) features, returned as fp32 like every other output
            ref = ref.float()
        if self._interleaved(k, ref):
            return ref.index_select(3, self._inverse_perm(k)).permute(0, 3, 1, 2)
        return ref.permute(0, 3, 1, 2)

    def _inverse_perm(self, k: int) -> Tensor:
        cache = self.__dict__.setdefault("_inv_perm", {})
        if k not in cache:
            perm = self.perms[k]
            inv = [0] * len(perm)
            for pos, ch in enumerate(perm):
                inv[ch] = pos
            cache[k] = torch.tensor(inv, dtype=torch.long, device=self.device)
        return cache[k]

    def run_cascade(self, net, feats: List[List[Tensor]], proj_matrices: Dict[str, Tensor], depth_values: Tensor,
                    attn_temp: Optional[float] = None, shard=None) -> Dict:
        """feats[k][i]: NHWC features at stage k; i = 0 is the reference view, i >= 1 the source
        views (all of them, or with ``shard`` this rank's ``shard.views`` in order)."""
        temp = float(net.stagenet.attn_temp if attn_temp is None else attn_temp)
        outputs: Dict = {}
        prev = None
        dv = depth_values.to(device=self.device, dtype=torch.float32).contiguous()
        for p, wts in zip(self.plans, self.stage_weights):
            prev = self._run_stage(net, p, wts, feats[p.k], proj_matrices, dv, prev, temp, shard)
            outputs[f"stage{p.k + 1}"] = prev
            outputs.update(prev)
        return outputs

    # ------------------------------------------------------------------ two-stream forward
    def _forward_overlapped(self, net, x: Tensor, B: int, n_own: int, proj_matrices: Dict[str, Tensor], depth_values: Tensor,
                            npass: int, gen: int, shard=None) -> Dict:
        """The early cascade stages work on 1/64, 1/16 and 1/4 of the last stage's voxels: ~45 short, latency-bound launches
        (0.7 ms of 2.1 ms at cfg2) that leave most SMs idle, and stage k only needs pyramid level k.  So they run on a second,
        high-priority stream next to the pyramid's remaining large layers: level events fork the side stream, the persistent
        convolution kernels of the two streams split the SMs through mvster_set_sm_budget, and the last stage joins on the main
        stream.  Works eagerly and under CUDA-graph capture (fork/join through events)."""
        lib = _lib.load()
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)
            self._sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        side = self._side
        # SM split between the persistent convolution launches of the two streams while they overlap: the main stream's leave
        # `reserve` SMs free, the side stream's claim at most `side_sms` (0 = no limit).  Measured at cfg2: see DESIGN.md 3.5.
        reserve, side_sms = int(os.environ.get("MVSTER_MAIN_RESERVE", "48")), int(os.environ.get("MVSTER_SIDE_SMS", "0"))
        events = [torch.cuda.Event() for _ in range(4)]

        stored: List[Optional[Tensor]] = [None] * 4
        as_bf16 = bf16_storage(net)

        def on_level(k, t):
            if as_bf16:  # storage='bf16': the level as it is stored, cast on the main stream before the cascade stage may start
                stored[k] = capi.cast_bf16(t)
            events[k].record(main)
            if k == 0 and reserve > 0:
                lib.mvster_set_sm_budget(self._sm_count - reserve)

        temp = float(net.stagenet.attn_temp)
        dv = depth_values.to(device=self.device, dtype=torch.float32).contiguous()
        outputs: Dict = {}
        deferred: List = []
        try:
            # relative poses depend on the inputs only: computed up front, off the stage-to-stage chain - and off the main stream,
            # whose first launch is the pyramid's stem (the four one-warp pose launches were 18 us in front of it)
            pose_kw = {} if shard is None else dict(first_view=shard.first_view, n_views=shard.count)  # a view shard: its own source views only
            projs = [proj_matrices[f"stage{k + 1}"].to(device=self.device, dtype=torch.float32).contiguous() for k in range(4)]
            fork = torch.cuda.Event()
            fork.record(main)
            with torch.cuda.stream(side):
                side.wait_event(fork)  # the inputs were produced on (or before) the main stream
                poses = [capi.pose(pr, **pose_kw) for pr in projs]
            pyramid = fpn_engine.run_fpn(self.fpn_weights, x, npass, gen=gen, on_level=on_level)
            levels = stored if as_bf16 else [pyramid[f"stage{k + 1}"] for k in range(4)]
            feats = [[levels[k][i * B:(i + 1) * B] for i in range(n_own)] for k in range(4)]
            # outputs that only depend on the pyramid go out here, on the main stream, in the gap between the pyramid's last launch
            # and the join with the side stream (at the end of the forward they were 37 us of serialised tail)
            mono = [self._mono_feat(k, feats[k][0]) if net.mono else None for k in range(4)]
            lib.mvster_set_sm_budget(side_sms)
            prev = None
            with torch.cuda.stream(side):
                for p, wts in zip(self.plans[:3], self.stage_weights[:3]):
                    side.wait_event(events[p.k])
                    prev = self._run_stage(net, p, wts, feats[p.k], proj_matrices, dv, prev, temp, shard=shard, pose=poses[p.k], deferred=deferred,
                                           mono_feat=mono[p.k])
                    outputs[f"stage{p.k + 1}"] = prev
                done = torch.cuda.Event()
                done.record(side)
                for fn in deferred:  # confidence maps of stages 1-3 (outputs only): on the side stream, next to stage 4
                    fn()
                tail = torch.cuda.Event()
                tail.record(side)
            lib.mvster_set_sm_budget(0)
            main.wait_event(done)
            prev = self._run_stage(net, self.plans[3], self.stage_weights[3], feats[3], proj_matrices, dv, prev, temp, shard=shard, pose=poses[3],
                                   mono_feat=mono[3])
            main.wait_event(tail)  # join: everything the forward returns is ordered before the caller's stream continues
            outputs["stage4"] = prev
            outputs.update(prev)
        finally:
            lib.mvster_set_sm_budget(0)
        return outputs
