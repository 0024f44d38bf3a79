"""Geometric-consistency filtering of depth maps on the GPU - the step that follows the forward in the reference's
``test_mvs4.py`` (``reproject_with_depth`` :271-309, ``check_geometric_consistency`` :312-328, the per-reference-view body of
``filter_depth`` :348-382), there numpy + ``cv2.remap`` in a process pool.  SURVEY.md 8f "next" #3.

Same signatures and return values as the reference's functions (numpy in, numpy out by default; CUDA tensors in, tensors out),
one launch of ``mvster_geo_consistency_f32`` (csrc/fusion.cu) per (reference, source) pair.  The camera matrices are prepared
on the host with the reference's own numpy calls - float32 inverses and products - so the kernel starts from bit-identical
matrices.  No CPU fallback: without the library the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib, capi

Tensor = torch.Tensor
Array = Union[np.ndarray, Tensor]


def remap_bilinear_reference(src: np.ndarray, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """numpy statement of what the kernel's source-depth lookup computes = ``cv2.remap(src, x, y, INTER_LINEAR)`` with the
    default constant-zero border (bit-exact, tests/test_fusion_oracle.py): documentation and test helper, not used by the
    product path."""
    H, W = src.shape
    f32 = np.float32
    sx = np.rint(x.astype(f32) * f32(32)).astype(np.int64)
    sy = np.rint(y.astype(f32) * f32(32)).astype(np.int64)
    ix, iy = sx >> 5, sy >> 5
    fx, fy = (sx & 31).astype(f32) * f32(0.03125), (sy & 31).astype(f32) * f32(0.03125)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        return np.where(ok, src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], f32(0))

    ex, ey = f32(1) - fx, f32(1) - fy
    r = tap(iy, ix) * (ey * ex)
    r = r + tap(iy, ix + 1) * (ey * fx)
    r = r + tap(iy + 1, ix) * (fy * ex)
    return (r + tap(iy + 1, ix + 1) * (fy * fx)).astype(f32)


def _np32(m) -> np.ndarray:
    return (m.detach().cpu().numpy() if isinstance(m, Tensor) else np.asarray(m)).astype(np.float32, copy=False)


def _prepare_mats(k_ref, e_ref, k_src, e_src) -> np.ndarray:
    """The six matrices of the reprojection chain, computed exactly as test_mvs4.py:279-303 computes them (float32 numpy),
    flattened as doubles: inv(K_ref) | (E_src inv(E_ref))[:3] | K_src | inv(K_src) | (E_ref inv(E_src))[:3] | K_ref."""
    k_ref, e_ref, k_src, e_src = _np32(k_ref)[:3, :3], _np32(e_ref), _np32(k_src)[:3, :3], _np32(e_src)
    parts = [np.linalg.inv(k_ref), np.matmul(e_src, np.linalg.inv(e_ref))[:3], k_src, np.linalg.inv(k_src),
             np.matmul(e_ref, np.linalg.inv(e_src))[:3], k_ref]
    return np.concatenate([p.astype(np.float64).reshape(-1) for p in parts])


def _dev_depth(d: Array, device) -> Tensor:
    t = d if isinstance(d, Tensor) else torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def _launch(depth_ref: Tensor, depth_src: Tensor, mats: np.ndarray, want_xy: bool, want_mask: bool,
            mask_sum: Tensor = None, depth_sum: Tensor = None, dist_thres: float = 1.0, rel_thres: float = 0.01):
    H, W = depth_ref.shape
    Hs, Ws = depth_src.shape
    dev = depth_ref.device
    out = torch.empty((H, W), device=dev, dtype=torch.float32)
    xs = torch.empty((H, W), device=dev, dtype=torch.float32) if want_xy else None
    ys = torch.empty((H, W), device=dev, dtype=torch.float32) if want_xy else None
    mask = torch.empty((H, W), device=dev, dtype=torch.uint8) if want_mask else None
    m = (C.c_double * 60)(*mats.tolist())
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mvster_geo_consistency_f32(capi._ptr(depth_ref), capi._ptr(depth_src), m, capi._ptr(out), capi._ptr(xs),
                                                          capi._ptr(ys), capi._ptr(mask), capi._ptr(mask_sum), capi._ptr(depth_sum),
                                                          H, W, Hs, Ws, float(dist_thres), float(rel_thres), capi._stream()),
                   "mvster_geo_consistency_f32")
    return out, xs, ys, mask


def check_geometric_consistency(depth_ref: Array, intrinsics_ref, extrinsics_ref, depth_src: Array, intrinsics_src, extrinsics_src,
                                device=None):
    """``test_mvs4.py:312-328``: (mask, depth_reprojected with rejected pixels zeroed, x2d_src, y2d_src).  numpy inputs give numpy
    outputs (mask bool), CUDA tensors give tensors."""
    as_numpy = not isinstance(depth_ref, Tensor)
    device = torch.device(device) if device is not None else (depth_ref.device if not as_numpy else torch.device("cuda", torch.cuda.current_device()))
    dr, ds = _dev_depth(depth_ref, device), _dev_depth(depth_src, device)
    out, xs, ys, mask = _launch(dr, ds, _prepare_mats(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src), True, True)
    if as_numpy:
        return mask.cpu().numpy().astype(bool), out.cpu().numpy(), xs.cpu().numpy(), ys.cpu().numpy()
    return mask.bool(), out, xs, ys


def fuse_reference_view(ref_depth: Array, ref_intrinsics, ref_extrinsics, confidence: Array,
                        sources: Sequence[Tuple[Array, np.ndarray, np.ndarray]], conf_thres: float, thres_view: int, device=None) -> Dict:
    """The per-reference-view body of ``filter_depth`` (``test_mvs4.py:348-382``): photometric mask ``confidence > conf_thres``,
    number of source views that pass the geometric check per pixel, ``geo_mask = count >= thres_view``, their conjunction, and the
    averaged depth ``(sum of accepted reprojections + reference) / (count + 1)`` (float64 like the reference).  The count and the
    float32 sum are accumulated inside the kernel launches, one per source view, in the order given."""
    as_numpy = not isinstance(ref_depth, Tensor)
    device = torch.device(device) if device is not None else (ref_depth.device if not as_numpy else torch.device("cuda", torch.cuda.current_device()))
    dr = _dev_depth(ref_depth, device)
    conf = _dev_depth(confidence, device)
    count = torch.zeros(dr.shape, device=device, dtype=torch.int32)
    dsum = torch.zeros(dr.shape, device=device, dtype=torch.float32)
    for src_depth, src_k, src_e in sources:
        _launch(dr, _dev_depth(src_depth, device), _prepare_mats(ref_intrinsics, ref_extrinsics, src_k, src_e), False, False, count, dsum)
    averaged = (dsum + dr).double() / (count + 1).double()
    photo = conf > conf_thres
    geo = count >= thres_view
    res = {"photo_mask": photo, "geo_mask": geo, "final_mask": photo & geo, "geo_mask_sum": count, "depth_est_averaged": averaged}
    return {k: v.cpu().numpy() for k, v in res.items()} if as_numpy else res


def backproject_points(depth: Array, mask: Array, intrinsics, extrinsics, image: Array = None):
    """World-space points of the pixels selected by ``mask`` (``test_mvs4.py:384-399`` of ``filter_depth``): pixel (x, y) with depth
    d -> ``inv(K) [x, y, 1] d`` in the camera frame -> ``inv(E)`` to the world.  Arithmetic as in the reference (float64 products
    with float32 matrix inverses).  Returns (xyz [N,3] float32, rgb [N,3] uint8 or None); works on numpy arrays or tensors of any
    device and returns the same kind."""
    as_numpy = not isinstance(depth, Tensor)
    d = torch.as_tensor(np.asarray(depth)) if as_numpy else depth
    m = torch.as_tensor(np.asarray(mask)).to(d.device) if not isinstance(mask, Tensor) else mask.to(d.device)
    H, W = d.shape
    k_inv = torch.from_numpy(np.linalg.inv(_np32(intrinsics)[:3, :3]).astype(np.float64)).to(d.device)
    e_inv = torch.from_numpy(np.linalg.inv(_np32(extrinsics)).astype(np.float64)).to(d.device)
    ys, xs = torch.nonzero(m.bool(), as_tuple=True)          # row-major order = the reference's boolean indexing
    dv = d[ys, xs].to(torch.float64)
    pix = torch.stack((xs.to(torch.float64) * dv, ys.to(torch.float64) * dv, dv))      # [3,N]
    cam = k_inv @ pix
    world = (e_inv @ torch.cat((cam, torch.ones_like(cam[:1]))))[:3]
    xyz = world.t().to(torch.float32).contiguous()
    rgb = None
    if image is not None:
        img = torch.as_tensor(np.asarray(image)).to(d.device) if not isinstance(image, Tensor) else image.to(d.device)
        rgb = (img[ys, xs] * 255).to(torch.uint8)              # (color * 255).astype(np.uint8): truncation
    if as_numpy:
        return xyz.cpu().numpy(), (None if rgb is None else rgb.cpu().numpy())
    return xyz, rgb
