"""Training-side use of the fused warp + epipolar-Transformer kernels: an autograd node whose forward is
``mvster_et_fuse_f32`` and whose backward is ``mvster_et_fuse_bwd_f32`` (csrc/et_fuse_bwd.cu).

The reference differentiates ``stagenet.forward`` (mvs4net_utils.py:1015-1062) through PyTorch ops and therefore keeps, for
every source view of every stage, the warped volume [B,C,D,H,W], the correlation and the weights alive until the backward
pass (at 5 views 512x640, stage 4: 42 MB of warped volume, 21 MB of correlations and ~20 MB of grid and weights per source view).  Here the node saves the channels-last features it was given plus
the cost volume and the weight sum; the backward kernel gathers the taps again.  Gradients reach the features only: the
sampling grid is built under ``torch.no_grad()`` (:23) and the hypotheses are detached (MVS4Net.py:95).

There is no CPU implementation: without the CUDA library the calls raise (``_lib.load``).
"""
from __future__ import annotations

import os
from typing import Sequence

import torch
from torch import Tensor

from . import capi


def _on(device):
    """Make ``device`` current for the launches inside (they go to its current stream); a no-op for host tensors, which only
    the test emulation of the kernels passes."""
    import contextlib
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


def enabled_by_default() -> bool:
    """The training-mode aggregation of fp32 CUDA features runs through the fused kernel and its hand-written backward unless
    MVSTER_TRAIN_ET=0 (then: PyTorch ops).  Measured on B200 at the cfg2 stage shapes (profiles/r02_et_backward.md): forward +
    backward 0.29-0.53 ms against 2.9-3.6 ms, 4.2-6.7 x less memory held between the passes, gradients within 1.6e-4 of max."""
    return os.environ.get("MVSTER_TRAIN_ET", "1") == "1"


def usable(features: Sequence[Tensor], G: int = 0, D: int = 0) -> bool:
    """The kernels take fp32 CUDA feature maps in the shape classes mvster_et_fuse_bwd_f32 is instantiated for
    (G in {4, 8}, C/G in {1, 2, 4, 8}, D in {4, 8}; pass G / D to have them checked); anything else stays on the PyTorch
    formulation instead of raising from inside a training step."""
    if not all(f.is_cuda and f.dtype == torch.float32 for f in features):
        return False
    if G:
        C = features[0].shape[1]
        if G not in (4, 8) or C % G or C // G not in (1, 2, 4, 8):
            return False
    return D in (0, 4, 8)


class EtFuse(torch.autograd.Function):
    """cost [B,D,H,W,G] = ET aggregation of channels-last features; see ``capi.et_fuse`` / ``capi.et_fuse_bwd``."""

    @staticmethod
    def forward(ctx, pose: Tensor, hypo: Tensor, G: int, attn_temp: float, ref: Tensor, *srcs: Tensor) -> Tensor:
        B, H, W, _ = ref.shape
        D = hypo.shape[1]
        wsum = torch.empty((B, D, H, W), device=ref.device, dtype=torch.float32)
        with _on(ref.device):  # launch on the tensors' device and its current stream, whatever is current
            cost = capi.et_fuse(ref, srcs, pose, hypo, G, attn_temp, wsum=wsum, partial=True)
            capi.et_normalize(cost, wsum)
        ctx.save_for_backward(pose, hypo, cost, wsum, ref, *srcs)
        ctx.attn_temp = attn_temp
        return cost

    @staticmethod
    def backward(ctx, grad_cost: Tensor):
        pose, hypo, cost, wsum, ref, *srcs = ctx.saved_tensors
        need = ctx.needs_input_grad[4:]
        if not any(need):
            return (None,) * (4 + 1 + len(srcs))
        with _on(ref.device):
            grad_ref, grad_src = capi.et_fuse_bwd(ref, srcs, pose, hypo, cost, wsum, grad_cost.contiguous(), ctx.attn_temp,
                                                  need_src=need[1:])
        return (None, None, None, None, grad_ref if need[0] else None, *grad_src)


def aggregate(features: Sequence[Tensor], cams: Tensor, hypo: Tensor, G: int, attn_temp: float) -> Tensor:
    """Differentiable drop-in for ``torch_path.aggregate(..., group_cor=True, attn_fuse_d=True)``: features Nv x [B,C,H,W]
    (view 0 = reference), cams [B,Nv,2,4,4], hypo [B,D,H,W] -> cost [B,G,D,H,W] (a permuted view of the kernel's
    channels-last volume)."""
    with torch.no_grad(), _on(features[0].device):
        pose = capi.pose(cams.contiguous().float())
        hypo = hypo.detach().contiguous().float()
    nhwc = [f.permute(0, 2, 3, 1).contiguous() for f in features]
    cost = EtFuse.apply(pose, hypo, int(G), float(attn_temp), *nhwc)
    return cost.permute(0, 4, 1, 2, 3)


# ---------------------------------------------------------------------------------------------------------------------------
# optimal-transport loss (mvs4net_utils.py:1096-1142), forward and backward in one launch (csrc/sinkhorn.cu)
def sinkhorn_enabled() -> bool:
    """The OT term of MVS4net_loss / Blend_loss runs on the fused kernel unless MVSTER_TRAIN_SINKHORN=0 (then: PyTorch ops).  Measured
    on B200 at the cfg2 stage shapes, 10 iterations (profiles/r02_sinkhorn.md): forward + backward 0.17-0.22 ms per stage against
    2.9-4.1 ms; loss and gradients within 1e-4 of values taken through the unmodified reference."""
    return os.environ.get("MVSTER_TRAIN_SINKHORN", "1") == "1"


def sinkhorn_usable(attn: Tensor, iters: int) -> bool:
    return attn.is_cuda and attn.dtype == torch.float32 and attn.shape[1] in (4, 8) and 0 <= int(iters) <= 32


class SinkhornLoss(torch.autograd.Function):
    """Masked mean of the per-pixel transport cost; the gradient w.r.t. ``attn`` comes from the same launch."""

    @staticmethod
    def forward(ctx, attn: Tensor, gt: Tensor, hypo: Tensor, mask: Tensor, iters: int, eps: float, continuous: bool) -> Tensor:
        with _on(attn.device):
            loss_px, grad = capi.sinkhorn(gt.detach().contiguous().float(), hypo.detach().contiguous().float(), attn.detach().contiguous(),
                                          mask.contiguous(), iters, eps, continuous, want_grad=attn.requires_grad)
        count = mask.sum()
        ctx.save_for_backward(grad, count)
        return loss_px.sum() / count  # loss_px is 0 outside the mask; an empty mask gives NaN like the reference's mean of nothing

    @staticmethod
    def backward(ctx, g: Tensor):
        grad, count = ctx.saved_tensors
        return (None if grad is None else grad * (g / count), None, None, None, None, None, None)


def sinkhorn_loss(gt: Tensor, hypo: Tensor, attn: Tensor, mask: Tensor, iters: int, eps: float = 1.0, continuous: bool = False) -> Tensor:
    """Differentiable drop-in for ``losses.sinkhorn(...)[1]``."""
    return SinkhornLoss.apply(attn, gt, hypo, mask, int(iters), float(eps), bool(continuous))

