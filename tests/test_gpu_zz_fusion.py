"""Geometric-consistency filter on the GPU (mvster_b200/fusion.py -> mvster_geo_consistency_f32) against the reference's own
outputs (tests/golden/fusion/*.npz, produced by executing test_mvs4.py:271-328) and the CPU oracle.  Coordinates and depths must
agree to float32 rounding; masks may differ only on pixels that sit on a threshold (|p_reproj - p| = 1 px or
|d_reproj - d| / d = 1 %) to within the float64 summation-order noise of the matrix products."""
import numpy as np
import pytest
import torch

from util import GOLDEN
from oracle import fusion_oracle

from mvster_b200 import _lib, fusion

pytestmark = pytest.mark.gpu
CASES = ["plane_4v_48x64", "plane_3v_40x56_wide"]


def load(name):
    z = np.load(GOLDEN / "fusion" / f"{name}.npz")
    nv, H, W, seed = (int(x) for x in z["meta"])
    return z, fusion_oracle.synthetic_scene(nv, H, W, seed, float(z["step"]))


def on_threshold(ref, s):
    """pixels whose consistency test is decided by less than 1e-4 px / 1e-6 relative: the only ones allowed to flip"""
    dr, xr, yr, _, _ = fusion_oracle.reproject_with_depth(ref["depth"], ref["K"], ref["E"], s["depth"], s["K"], s["E"])
    H, W = ref["depth"].shape
    xx, yy = np.meshgrid(np.arange(W), np.arange(H))
    dist = np.sqrt((xr - xx) ** 2 + (yr - yy) ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(dr - ref["depth"]) / ref["depth"]
    return (np.abs(dist - 1) < 1e-4) | (np.abs(rel - 0.01) < 1e-6)


@pytest.mark.parametrize("name", CASES)
def test_check_geometric_consistency_matches_reference(name):
    z, views = load(name)
    ref = views[0]
    n0 = _lib.launch_count()
    for v in range(1, len(views)):
        s = views[v]
        mask, dr, xs, ys = fusion.check_geometric_consistency(ref["depth"], ref["K"], ref["E"], s["depth"], s["K"], s["E"])
        assert mask.dtype == bool and dr.dtype == np.float32 and xs.dtype == np.float32
        for got, want in ((xs, z[f"x2d_src{v}"]), (ys, z[f"y2d_src{v}"])):
            finite = np.isfinite(want)
            assert np.array_equal(np.isfinite(got), finite)
            ulp = np.spacing(np.abs(want[finite]).astype(np.float32))
            assert (np.abs(got[finite] - want[finite]) <= ulp).all()           # float32 cast of a float64 that differs in the last bits
            assert (got[finite] != want[finite]).mean() < 0.01
        flips = mask != z[f"mask{v}"]
        assert flips.mean() < 0.005 and not (flips & ~on_threshold(ref, s)).any(), f"{int(flips.sum())} mask flips off the thresholds"
        both = mask & z[f"mask{v}"]
        want_d = z[f"depth_reprojected{v}"]
        assert (np.abs(dr[both] - want_d[both]) <= 2e-6 * want_d[both]).all()
        assert not dr[~mask].any()                                              # rejected pixels are zeroed
    assert _lib.launch_count() - n0 == len(views) - 1


@pytest.mark.parametrize("name", CASES)
def test_fuse_reference_view_matches_reference(name):
    z, views = load(name)
    ref = views[0]
    fused = fusion.fuse_reference_view(ref["depth"], ref["K"], ref["E"], z["confidence"],
                                       [(s["depth"], s["K"], s["E"]) for s in views[1:]], 0.5, 2)
    edge = np.zeros(ref["depth"].shape, bool)
    for s in views[1:]:
        edge |= on_threshold(ref, s)
    same = fused["geo_mask_sum"] == z["geo_mask_sum"]
    assert (same | edge).all() and same.mean() > 0.99
    assert np.array_equal(fused["photo_mask"], z["confidence"] > 0.5)
    assert np.array_equal(fused["final_mask"][same], z["final_mask"][same])
    assert fused["depth_est_averaged"].dtype == np.float64
    a, b = fused["depth_est_averaged"][same], z["depth_est_averaged"][same]
    assert (np.abs(a - b) <= 2e-6 * np.abs(b)).all()


def test_identity_pair_and_cuda_tensor_interface():
    """A view checked against itself reprojects every pixel with a non-zero depth onto itself (the reference accepts negative
    depths as well: only d == 0 fails, through 0 / 0): same mask as the CPU oracle, depth unchanged to rounding; CUDA tensors in ->
    CUDA tensors out."""
    views = fusion_oracle.synthetic_scene(2, 96, 160, seed=3)
    v = views[0]
    v["depth"][5, 7] = 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        want_mask, want_d, _, _ = fusion_oracle.check_geometric_consistency(v["depth"].copy(), v["K"], v["E"], v["depth"].copy(), v["K"], v["E"])
    d = torch.from_numpy(v["depth"]).cuda()
    mask, dr, xs, ys = fusion.check_geometric_consistency(d, v["K"], v["E"], d, v["K"], v["E"])
    assert mask.is_cuda and mask.dtype == torch.bool and dr.is_cuda
    assert np.array_equal(mask.cpu().numpy(), want_mask) and not want_mask[5, 7] and want_mask.mean() > 0.99
    m = torch.from_numpy(want_mask).cuda()
    assert ((dr[m] - d[m]).abs() <= 1e-5 * d[m].abs()).all() and not dr[~m].any()
    assert (dr.cpu().numpy()[want_mask] == want_d[want_mask]).mean() > 0.99
    xx = torch.arange(160, device="cuda", dtype=torch.float32).expand(96, 160)
    assert ((xs - xx)[m].abs() < 1e-3).all()
