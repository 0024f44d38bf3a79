"""CPU restatement of the reference's geometric-consistency filter - TEST INFRASTRUCTURE ONLY (tests/, smoke and bench
baselines may import it; the product path in mvster_b200/ never does).

Follows test_mvs4.py of the reference line by line:
  reproject_with_depth          test_mvs4.py:271-309
  check_geometric_consistency   test_mvs4.py:312-328
  fuse_reference_view           the per-reference-view body of filter_depth, test_mvs4.py:348-382 (masks and averaged depth;
                                file I/O, mask PNGs and the PLY writer are not part of it)
Arithmetic as in the reference: numpy float64 geometry on float32 camera matrices, cv2.remap(INTER_LINEAR) for the source-depth
lookup (fixed-point coordinates, 1/32 pixel; zero outside the image).  Pinned by tests/golden/fusion/*.npz, which
oracle/make_fusion_golden.py produces by executing the reference's own function source.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import cv2
import numpy as np


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """test_mvs4.py:271-309.  Returns (depth_reprojected, x_reprojected, y_reprojected, x_src, y_src), all [H,W] float32."""
    width, height = depth_ref.shape[1], depth_ref.shape[0]
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), np.vstack((x_ref, y_ref, np.ones_like(x_ref))) * depth_ref.reshape([-1]))
    xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, np.ones_like(x_ref))))[:3]
    k_xyz_src = np.matmul(intrinsics_src, xyz_src)
    xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
    x_src = xy_src[0].reshape([height, width]).astype(np.float32)
    y_src = xy_src[1].reshape([height, width]).astype(np.float32)
    sampled_depth_src = cv2.remap(depth_src, x_src, y_src, interpolation=cv2.INTER_LINEAR)
    xyz_src = np.matmul(np.linalg.inv(intrinsics_src), np.vstack((xy_src, np.ones_like(x_ref))) * sampled_depth_src.reshape([-1]))
    xyz_reprojected = np.matmul(np.matmul(extrinsics_ref, np.linalg.inv(extrinsics_src)), np.vstack((xyz_src, np.ones_like(x_ref))))[:3]
    depth_reprojected = xyz_reprojected[2].reshape([height, width]).astype(np.float32)
    k_xyz_reprojected = np.matmul(intrinsics_ref, xyz_reprojected)
    xy_reprojected = k_xyz_reprojected[:2] / k_xyz_reprojected[2:3]
    x_reprojected = xy_reprojected[0].reshape([height, width]).astype(np.float32)
    y_reprojected = xy_reprojected[1].reshape([height, width]).astype(np.float32)
    return depth_reprojected, x_reprojected, y_reprojected, x_src, y_src


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """test_mvs4.py:312-328: reprojection error < 1 px and relative depth difference < 1 %.  Returns (mask, depth_reprojected
    with the rejected pixels zeroed, x2d_src, y2d_src)."""
    width, height = depth_ref.shape[1], depth_ref.shape[0]
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_reprojected, x2d_reprojected, y2d_reprojected, x2d_src, y2d_src = reproject_with_depth(
        depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
    dist = np.sqrt((x2d_reprojected - x_ref) ** 2 + (y2d_reprojected - y_ref) ** 2)
    depth_diff = np.abs(depth_reprojected - depth_ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        relative_depth_diff = depth_diff / depth_ref
    mask = np.logical_and(dist < 1, relative_depth_diff < 0.01)
    depth_reprojected[~mask] = 0
    return mask, depth_reprojected, x2d_src, y2d_src


def fuse_reference_view(ref_depth, ref_intrinsics, ref_extrinsics, confidence,
                        sources: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], conf_thres: float, thres_view: int) -> Dict:
    """test_mvs4.py:348-382.  sources = [(depth_est, intrinsics, extrinsics)] of the source views.  Returns the photometric,
    geometric and final masks, the per-pixel count of consistent source views and the averaged depth (float64, as in the
    reference: float32 sum / int32 count)."""
    photo_mask = confidence > conf_thres
    all_depth, geo_mask_sum = [], 0
    for src_depth, src_k, src_e in sources:
        geo_mask, depth_reprojected, _, _ = check_geometric_consistency(ref_depth, ref_intrinsics, ref_extrinsics, src_depth, src_k, src_e)
        geo_mask_sum = geo_mask_sum + geo_mask.astype(np.int32)
        all_depth.append(depth_reprojected)
    depth_est_averaged = (sum(all_depth) + ref_depth) / (geo_mask_sum + 1)
    geo_mask = geo_mask_sum >= thres_view
    return {"photo_mask": photo_mask, "geo_mask": geo_mask, "final_mask": np.logical_and(photo_mask, geo_mask),
            "geo_mask_sum": geo_mask_sum, "depth_est_averaged": depth_est_averaged}


def synthetic_scene(n_views: int, H: int, W: int, seed: int = 0, step_deg: float = 4.0) -> List[Dict[str, np.ndarray]]:
    """Depth maps of a tilted plane seen by cameras on an arc (mvster_b200.synth.arc_cameras), with a band of wrong depth and a
    few zeros in every view so that the consistency masks are mixed.  Deterministic; used by the golden script and the tests."""
    from mvster_b200 import synth
    rng = np.random.RandomState(seed)
    cams = synth.arc_cameras(n_views, H, W, step_deg)
    n, c = np.array([0.15, -0.1, 1.0]), 650.0  # plane n . X = c in world coordinates
    views = []
    for v in range(n_views):
        E, K = cams[v, 0].astype(np.float32), cams[v, 1, :3, :3].astype(np.float32)
        R, t = E[:3, :3].astype(np.float64), E[:3, 3].astype(np.float64)
        ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        rays = np.linalg.inv(K.astype(np.float64)) @ np.stack([xs.ravel(), ys.ravel(), np.ones(H * W)])  # camera frame, z = 1
        # world point = R^T (z * ray - t);  n . world = c  ->  z = (c + n . R^T t) / (n . R^T ray)
        nr = n @ R.T
        z = (c + nr @ t) / (nr @ rays)
        depth = z.reshape(H, W).astype(np.float32)
        depth[H // 3: H // 3 + 4, :] *= np.float32(1.0 + 0.004 * (v + 1))   # up to 2 % off: straddles the 1 % test
        depth[rng.randint(0, H, 6), rng.randint(0, W, 6)] = 0.0               # holes
        depth += (rng.rand(H, W).astype(np.float32) - 0.5) * np.float32(0.4)  # +-0.03 % noise
        views.append({"depth": depth, "K": K, "E": E})
    return views
