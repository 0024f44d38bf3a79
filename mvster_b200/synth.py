"""Seeded synthetic inputs for the MVSTER forward path (no dataset, no checkpoint).

Mirrors the tensor contract the reference's datasets hand to ``MVS4net.forward``
(datasets/dtu_yao4.py:167-195): a list of Nv images ``[B,3,H,W]`` in [0,1), a dict
``stage1..4 -> [B,Nv,2,4,4]`` of (extrinsic, intrinsic) pairs whose intrinsics are
pre-scaled by 1/8, 1/4, 1/2, 1, and ``depth_values [B,2]``.  View 0 is the
reference view.  Everything is produced by ``numpy.random.RandomState`` so the
numbers are identical on every machine.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np
import torch

DTU_DEPTH_RANGE = (425.0, 935.0)  # datasets/dtu_yao4.py:60-61,163: 425 + 2.5*1.06*192


def arc_cameras(n_views: int, H: int, W: int, step_deg: float = 1.0, radius: float = 680.0) -> np.ndarray:
    """Cameras on an arc looking at a point `radius` in front of the reference camera.
    Returns [Nv,2,4,4] float32 at FULL resolution (slot 0 extrinsic, slot 1 intrinsic)."""
    cams = np.zeros((n_views, 2, 4, 4), np.float32)
    K = np.array([[0.9 * W, 0, W / 2.0], [0, 0.9 * W, H / 2.0], [0, 0, 1]], np.float64)
    for v in range(n_views):
        if v == 0:
            R, c = np.eye(3), np.zeros(3)
        else:
            a = math.radians(step_deg * math.ceil(v / 2) * (1 if v % 2 else -1))
            R = np.array([[math.cos(a), 0, -math.sin(a)], [0, 1, 0], [math.sin(a), 0, math.cos(a)]])
            c = np.array([radius * math.sin(a), 3.0 * v, radius - radius * math.cos(a)])
        E = np.eye(4)
        E[:3, :3] = R
        E[:3, 3] = -R @ c
        cams[v, 0] = E.astype(np.float32)
        cams[v, 1, :3, :3] = K.astype(np.float32)
    return cams


def stage_projections(cams_full: np.ndarray, batch: int, num_stage: int = 4) -> Dict[str, torch.Tensor]:
    """Scale intrinsic rows 0-1 by 2^-(num_stage-1-k) for stage k (datasets/dtu_yao4.py:176-189)."""
    out = {}
    for k in range(num_stage):
        c = cams_full.copy()
        c[:, 1, :2, :] *= 1.0 / (2 ** (num_stage - 1 - k))
        out[f"stage{k + 1}"] = torch.from_numpy(np.repeat(c[None], batch, 0).copy())
    return out


def make_inputs(batch: int, n_views: int, H: int, W: int, seed: int = 0, step_deg: float = 1.0,
                depth_range: Tuple[float, float] = DTU_DEPTH_RANGE):
    """-> (imgs list of Nv [B,3,H,W], proj dict, depth_values [B,2]) as CPU fp32 tensors."""
    rng = np.random.RandomState(seed)
    imgs = [torch.from_numpy(rng.rand(batch, 3, H, W).astype(np.float32)) for _ in range(n_views)]
    proj = stage_projections(arc_cameras(n_views, H, W, step_deg), batch)
    dv = torch.tensor([list(depth_range)] * batch, dtype=torch.float32)
    return imgs, proj, dv


def make_features(batch: int, n_views: int, H: int, W: int, seed: int = 0,
                  channels: Sequence[int] = (64, 32, 16, 8)) -> List[Dict[str, torch.Tensor]]:
    """Synthetic N-view H x W x C feature pyramids (NCHW tensors, N(0,1)) for driving the
    stage cascade without the feature net."""
    rng = np.random.RandomState(seed + 7919)
    feats = []
    for _ in range(n_views):
        f = {}
        for k, c in enumerate(channels):
            s = 2 ** (len(channels) - 1 - k)
            f[f"stage{k + 1}"] = torch.from_numpy(rng.randn(batch, c, H // s, W // s).astype(np.float32))
        feats.append(f)
    return feats


def synthetic_state(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic weights for every key of an MVS4net ``state_dict`` (key -> shape).
    Conv kernels ~ U(+-sqrt(6/fan_in)); BN gamma, running_var ~ U(0.5,1.5); BN beta,
    running_mean, conv biases ~ N(0,0.1) so that BN folding is exercised."""
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        rng = np.random.RandomState((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if leaf == "running_var" or (leaf == "weight" and len(shape) == 1):
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf in ("running_mean", "bias"):
            a = rng.normal(0.0, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            if "conv7.0" in key or "conv9.0" in key or "conv11.0" in key:  # ConvTranspose: [Cin,Cout,...]
                fan_in = int(shape[0] * np.prod(shape[2:])) // 4 or 1
            b = math.sqrt(6.0 / fan_in)
            a = rng.uniform(-b, b, shape)
        out[key] = torch.from_numpy(np.asarray(a, np.float32).reshape(shape))
    return out
