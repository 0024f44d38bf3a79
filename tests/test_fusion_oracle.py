"""oracle/fusion_oracle.py (restatement of test_mvs4.py:271-382) against vectors produced by executing the reference's own
function source (oracle/make_fusion_golden.py), plus the cv2.remap emulation the CUDA kernel implements."""
import numpy as np
import pytest

from util import GOLDEN
from oracle import fusion_oracle

CASES = ["plane_4v_48x64", "plane_3v_40x56_wide"]


def load(name):
    z = np.load(GOLDEN / "fusion" / f"{name}.npz")
    nv, H, W, seed = (int(x) for x in z["meta"])
    return z, fusion_oracle.synthetic_scene(nv, H, W, seed, float(z["step"]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_outputs(name):
    z, views = load(name)
    ref = views[0]
    for v in range(1, len(views)):
        s = views[v]
        mask, dr, xs, ys = fusion_oracle.check_geometric_consistency(ref["depth"].copy(), ref["K"], ref["E"], s["depth"].copy(), s["K"], s["E"])
        assert np.array_equal(mask, z[f"mask{v}"])
        assert np.array_equal(dr, z[f"depth_reprojected{v}"])
        assert np.array_equal(xs, z[f"x2d_src{v}"]) and np.array_equal(ys, z[f"y2d_src{v}"])
    fused = fusion_oracle.fuse_reference_view(ref["depth"], ref["K"], ref["E"], z["confidence"],
                                              [(s["depth"], s["K"], s["E"]) for s in views[1:]], 0.5, 2)
    assert np.array_equal(fused["geo_mask_sum"], z["geo_mask_sum"])
    assert np.array_equal(fused["final_mask"], z["final_mask"])
    assert fused["depth_est_averaged"].dtype == np.float64
    assert np.array_equal(fused["depth_est_averaged"], z["depth_est_averaged"])


def test_remap_emulation_is_bit_exact_with_cv2():
    """The kernel's source-depth lookup: coordinates rounded to 1/32 pixel (round-half-even of x * 32 in float32), weights
    (1 - fy)(1 - fx) ... in float32, taps outside the image read as zero, sum ((t00 w00 + t01 w01) + t10 w10) + t11 w11."""
    import cv2
    from mvster_b200.fusion import remap_bilinear_reference
    rng = np.random.RandomState(0)
    src = (rng.rand(37, 53).astype(np.float32) * 500)
    x = (rng.rand(64, 80) * 63 - 5).astype(np.float32)
    y = (rng.rand(64, 80) * 47 - 5).astype(np.float32)
    x[0, :4] = [0.0, 52.0, 51.984375, -0.015625]
    want = cv2.remap(src, x, y, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(remap_bilinear_reference(src, x, y), want)
