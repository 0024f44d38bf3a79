"""Differentiable PyTorch formulation of one cascade stage - the TRAINING path.

The sm_100a kernels are forward-only; ``train_mvs4.py`` needs autograd through the stage
(gradients flow to the features and the regulariser weights, never to the sampling grid,
mvs4net_utils.py:23).  This module is what ``MVS4net.forward`` uses when ``self.training``
or gradients are enabled.  It runs on whatever device the tensors live on and is not an
inference fallback: inference without the CUDA library raises (see network.py).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def full_projection(cam: Tensor) -> Tensor:
    """[B,2,4,4] (E, K) -> 4x4 with K@E[:3,:4] on top (mvs4net_utils.py:1032-1035)."""
    full = cam[:, 0].clone()
    full[:, :3, :4] = cam[:, 1, :3, :3] @ cam[:, 0, :3, :4]
    return full


@torch.no_grad()
def sampling_grid(src_cam: Tensor, ref_cam: Tensor, hypo: Tensor, Hs: int, Ws: int) -> Tensor:
    """Normalised grid [B, D*H, W, 2] of the plane sweep (mvs4net_utils.py:23-45)."""
    B, D, H, W = hypo.shape
    P = full_projection(src_cam) @ torch.inverse(full_projection(ref_cam))
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=hypo.device),
                            torch.arange(W, dtype=torch.float32, device=hypo.device), indexing="ij")
    pix = torch.stack((xs.flatten(), ys.flatten(), torch.ones(H * W, device=hypo.device)))
    ray = P[:, :3, :3] @ pix.expand(B, 3, H * W)
    pts = ray[:, :, None, :] * hypo.reshape(B, 1, D, H * W) + P[:, :3, 3].reshape(B, 3, 1, 1)
    z = torch.where(pts[:, 2:3] == 0, torch.full_like(pts[:, 2:3], 1e-9), pts[:, 2:3])
    uv = pts[:, :2] / z
    grid = torch.stack((uv[:, 0] / ((Ws - 1) / 2) - 1, uv[:, 1] / ((Hs - 1) / 2) - 1), dim=-1)
    return grid.reshape(B, D * H, W, 2)


def aggregate(features: Sequence[Tensor], cams: Tensor, hypo: Tensor, group_cor: bool, G: int,
              attn_temp: float, attn_fuse_d: bool) -> Tensor:
    """Epipolar-Transformer aggregation (mvs4net_utils.py:1015-1062) -> [B,G',D,H,W]."""
    ref = features[0]
    B, C, H, W = ref.shape
    D = hypo.shape[1]
    ref_vol = ref[:, :, None].expand(B, C, D, H, W)
    wsum, acc = 1e-8, 0
    for v in range(1, len(features)):
        src = features[v]
        grid = sampling_grid(cams[:, v], cams[:, 0], hypo, src.shape[2], src.shape[3])
        warped = F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        warped = warped.reshape(B, C, D, H, W)
        if group_cor:
            cor = (warped * ref_vol).reshape(B, G, C // G, D, H, W).mean(2)
        else:
            cor = (ref_vol - warped) ** 2
        if attn_fuse_d:
            w = (torch.softmax(cor.sum(1) / attn_temp, 1) / math.sqrt(C))[:, None]
        else:
            w = torch.softmax(cor.sum(1), 1).amax(1)[:, None, None]
        wsum = wsum + w
        acc = acc + w * cor
    return acc / wsum


def head(logits: Tensor, hypo: Tensor, stage_idx: int, split_itv: float, inverse_depth: bool,
         training: bool) -> Dict[str, Tensor]:
    """mvs4net_utils.py:1066-1088."""
    attn = F.softmax(logits, dim=1)
    depth = torch.gather(hypo, 1, attn.argmax(1, keepdim=True)).squeeze(1)
    if training:
        conf = torch.tensor(0.0, dtype=torch.float32, device=logits.device)
    else:
        with torch.no_grad():
            conf = F.interpolate(attn.amax(1, keepdim=True), scale_factor=2 ** (3 - stage_idx), mode="bilinear",
                                 align_corners=True).squeeze(1)
    out = {"depth": depth, "photometric_confidence": conf, "hypo_depth": hypo, "attn_weight": attn}
    if inverse_depth:
        itv = 1.0 / hypo[:, 2] - 1.0 / hypo[:, 1]
        out["inverse_min_depth"] = 1 / depth + split_itv * itv
        out["inverse_max_depth"] = 1 / depth - split_itv * itv
    return out


def inverse_range_init(depth_values: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:71-77."""
    lo, hi = 1.0 / depth_values[:, -1], 1.0 / depth_values[:, 0]
    ramp = torch.arange(D, device=depth_values.device, dtype=depth_values.dtype).view(1, D, 1, 1).repeat(1, 1, H, W) / (D - 1)
    return 1.0 / (lo[:, None, None, None] + (hi - lo)[:, None, None, None] * ramp)


def inverse_range_schedule(inv_min: Tensor, inv_max: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:79-86."""
    ramp = torch.arange(D, device=inv_min.device, dtype=inv_min.dtype).view(1, D, 1, 1).repeat(1, 1, H // 2, W // 2) / (D - 1)
    inv = inv_max[:, None] + (inv_min - inv_max)[:, None] * ramp
    inv = F.interpolate(inv[:, None], [D, H, W], mode="trilinear", align_corners=True)[:, 0]
    return 1.0 / inv


def linear_range_init(depth_values: Tensor, D: int, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:61-69."""
    lo = depth_values[:, 0]
    step = (depth_values[:, -1] - lo) / (D - 1)
    s = lo[:, None] + torch.arange(D, device=lo.device, dtype=lo.dtype)[None] * step[:, None]
    return s[:, :, None, None].repeat(1, 1, H, W)


def linear_range_schedule(depth: Tensor, D: int, itv: Tensor, H: int, W: int) -> Tensor:
    """mvs4net_utils.py:88-99; itv [B]."""
    lo = depth - D / 2 * itv[:, None, None]
    hi = depth + D / 2 * itv[:, None, None]
    step = (hi - lo) / (D - 1)
    s = lo[:, None] + torch.arange(D, device=depth.device, dtype=depth.dtype).view(1, D, 1, 1) * step[:, None]
    return F.interpolate(s[:, None], [D, H, W], mode="trilinear", align_corners=True)[:, 0]
