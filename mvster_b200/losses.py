"""Training losses with the reference's signatures (models/MVS4Net.py:113-206) - plain PyTorch
(training-only; they read ``depth``, ``hypo_depth``, ``attn_weight`` and ``mono_depth`` from the per-stage output
dicts).  The optimal-transport term of fp32 CUDA tensors runs on the fused kernel of csrc/sinkhorn.cu
(mvster_b200/train_ops.py; MVSTER_TRAIN_SINKHORN=0 keeps it on PyTorch ops)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def sinkhorn(gt_depth: Tensor, hypo_depth: Tensor, attn_weight: Tensor, mask: Tensor, iters: int, eps: float = 1,
             continuous: bool = False):
    """Entropy-regularised optimal transport between the predicted distribution over the D
    hypotheses and the ground-truth bin, in the log domain (mvs4net_utils.py:1096-1142).
    Returns (transport map, masked mean transport cost)."""
    B, D, H, W = attn_weight.shape
    dev = gt_depth.device
    bins = torch.arange(D, dtype=torch.float32, device=dev)
    ground = (bins[None, :] - bins[:, None]).abs()  # |i - j|
    if not continuous:
        cost = ground[None, None].repeat(B, H * W, 1, 1)
        gt_bin = (hypo_depth - gt_depth[:, None]).abs().argmin(1).reshape(B * H * W, 1)
        mu = torch.zeros(B * H * W, D, dtype=hypo_depth.dtype, device=dev)
        mu.scatter_add_(1, gt_bin, torch.ones_like(mu[:, :1]))
        mu = mu.reshape(B, H * W, D)
    else:
        mu = torch.zeros((B, H * W, D + 1), dtype=torch.float32, device=dev)
        mu[:, :, -1] = 1
        cost = torch.zeros((B, H, W, D, D + 1), dtype=torch.float32, device=dev)
        cost[..., :D] = ground
        itv = 1 / hypo_depth[:, 2] - 1 / hypo_depth[:, 1]
        dist = (1 / gt_depth - 1 / hypo_depth[:, 0]) / itv
        dist[~mask] = 10
        cost[..., -1] = (dist[..., None] - bins).abs()
        cost = cost.reshape(B, H * W, D, D + 1)
    nu = attn_weight.permute(0, 2, 3, 1).reshape(B, H * W, D)
    log_mu, log_nu = (mu + 1e-12).log(), (nu + 1e-12).log()
    u, v = torch.zeros_like(log_nu), torch.zeros_like(log_mu)
    for _ in range(iters):
        v = log_mu - torch.logsumexp(cost / eps + u.unsqueeze(3), dim=2)
        u = log_nu - torch.logsumexp(cost / eps + v.unsqueeze(2), dim=3)
    T = (cost / eps + u.unsqueeze(3) + v.unsqueeze(2)).exp()
    loss = (T * cost).reshape(B * H * W, -1)[mask.reshape(-1)].sum(-1).mean()
    return T, loss


def _stage_terms(inputs, depth_gt_ms, mask_ms, kw):
    lw = kw.get("stage_lw", [1, 1, 1, 1])
    l1w, otw = kw.get("l1ot_lw", [0, 1])
    inverse, mono = kw.get("inverse_depth", False), kw.get("mono", False)
    dev = mask_ms["stage1"].device
    total = torch.tensor(0.0, dtype=torch.float32, device=dev)
    l1s, ots, oor = [], [], []
    last = None
    for idx, key in enumerate(k for k in inputs.keys() if "stage" in k):
        st = inputs[key]
        hypo, attn = st["hypo_depth"], st["attn_weight"]
        mask = mask_ms[key] > 0.5
        gt = depth_gt_ms[key]
        if mono and idx != 0:
            l1 = F.l1_loss(st["mono_depth"][mask], gt[mask], reduction="mean")
        else:
            l1 = torch.tensor(0.0, dtype=torch.float32, device=dev)
        if inverse:
            itv = (1 / hypo[:, 2] - 1 / hypo[:, 1]).abs()
            miss = ((1 / hypo - 1 / gt.unsqueeze(1)).abs() <= itv.unsqueeze(1)).sum(1) == 0
        else:
            itv = (hypo[:, 2] - hypo[:, 1]).abs()
            miss = ((hypo - gt.unsqueeze(1)).abs() <= itv.unsqueeze(1)).sum(1) == 0
        oor.append(miss[mask].float().mean())
        ot_args = dict(iters=kw.get("ot_iter", 3), eps=kw.get("ot_eps", 1), continuous=kw.get("ot_continous", False))
        from . import train_ops  # the fused forward + backward kernel for fp32 CUDA tensors (MVSTER_TRAIN_SINKHORN=0: PyTorch ops)
        if train_ops.sinkhorn_enabled() and train_ops.sinkhorn_usable(attn, ot_args["iters"]):
            ot = train_ops.sinkhorn_loss(gt, hypo, attn, mask, **ot_args)
        else:
            ot = sinkhorn(gt, hypo, attn, mask, **ot_args)[1]
        l1s.append(l1)
        ots.append(ot)
        total = total + lw[idx] * (l1w * l1 + otw * ot)
        last = (st["depth"], gt, mask)
    return total, l1s, ots, oor, last


def MVS4net_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """models/MVS4Net.py:113-155 -> (total, [l1]*S, [ot]*S, [out-of-range ratio]*S)."""
    total, l1s, ots, oor, _ = _stage_terms(inputs, depth_gt_ms, mask_ms, kwargs)
    return total, l1s, ots, oor


def Blend_loss(inputs, depth_gt_ms, mask_ms, **kwargs):
    """models/MVS4Net.py:158-206: adds the normalised end-point error metrics of the last stage."""
    total, l1s, ots, oor, (pred, gt, mask) = _stage_terms(inputs, depth_gt_ms, mask_ms, kwargs)
    span = (kwargs.get("depth_max", 100) - kwargs.get("depth_min", 1))[:, None, None]
    err = (pred * 128 / span - gt * 128 / span)[mask].abs()
    return total, l1s, ots, oor, err.mean(), (err <= 3).float().mean() * 100, (err <= 1).float().mean() * 100
