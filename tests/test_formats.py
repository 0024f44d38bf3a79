"""mvster_b200/formats.py (PFM, camera files, pair lists, per-stage projections) against fixtures written / parsed by the
unmodified reference (tests/golden/make_format_fixtures.py: datasets/data_io.py:6-71, datasets/general_eval4.py:25-79,
155-183), and - when the reference tree is mounted - against the reference's functions directly."""
import json
import os
import sys

import numpy as np
import pytest

from util import GOLDEN
from mvster_b200 import formats

FIX = GOLDEN / "formats"


@pytest.mark.parametrize("name,scale", [("depth_6x9", 1.0), ("color_4x5", 2.0), ("big_endian_3x4", 1.0)])
def test_read_pfm_matches_reference_written_files(name, scale):
    data, s = formats.read_pfm(str(FIX / f"{name}.pfm"))
    want = np.load(FIX / f"{name}.npy")
    assert s == scale
    assert data.shape == want.shape
    assert np.array_equal(np.asarray(data, np.float32), want)  # bit-exact


@pytest.mark.parametrize("name,scale", [("depth_6x9", 1), ("color_4x5", 2)])
def test_save_pfm_is_byte_identical_to_the_reference_writer(tmp_path, name, scale):
    img = np.load(FIX / f"{name}.npy")
    p = tmp_path / "out.pfm"
    formats.save_pfm(str(p), img, scale=scale)
    assert p.read_bytes() == (FIX / f"{name}.pfm").read_bytes()
    back, s = formats.read_pfm(str(p))
    assert s == float(scale) and np.array_equal(np.asarray(back, np.float32), img)


def test_save_pfm_rejects_what_the_reference_rejects(tmp_path):
    with pytest.raises(Exception, match="float32"):
        formats.save_pfm(str(tmp_path / "a.pfm"), np.zeros((2, 2), np.float64))
    with pytest.raises(Exception, match="dimensions"):
        formats.save_pfm(str(tmp_path / "a.pfm"), np.zeros((2, 2, 2), np.float32))
    (tmp_path / "bad.pfm").write_bytes(b"P6\n2 2\n-1.0\n")
    with pytest.raises(Exception, match="Not a PFM"):
        formats.read_pfm(str(tmp_path / "bad.pfm"))
    (tmp_path / "bad2.pfm").write_bytes(b"Pf\n2x2\n-1.0\n")
    with pytest.raises(Exception, match="Malformed"):
        formats.read_pfm(str(tmp_path / "bad2.pfm"))


def test_single_channel_3d_image_is_written_as_greyscale(tmp_path):
    img = np.arange(6, dtype=np.float32).reshape(2, 3, 1)
    formats.save_pfm(str(tmp_path / "g.pfm"), img)
    back, _ = formats.read_pfm(str(tmp_path / "g.pfm"))
    assert back.shape == (2, 3) and np.array_equal(back, img[:, :, 0])


def test_read_cam_file_matches_the_reference_reader():
    expected = json.loads((FIX / "cam_expected.json").read_text())
    for key, want in expected.items():
        name, scale, nd = key.split("|")
        k, e, dmin, itv = formats.read_cam_file(str(FIX / name), float(scale), int(nd))
        assert np.array_equal(k, np.array(want["K"], np.float32)), key
        assert np.array_equal(e, np.array(want["E"], np.float32)), key
        assert dmin == want["depth_min"] and itv == want["depth_interval"], key


def test_cam_file_round_trip(tmp_path):
    k, e, dmin, itv = formats.read_cam_file(str(FIX / "00000000_cam.txt"))
    k_full = k.copy()
    k_full[:2] *= 4.0
    formats.write_cam_file(str(tmp_path / "c.txt"), e, k_full, dmin, itv)
    k2, e2, dmin2, itv2 = formats.read_cam_file(str(tmp_path / "c.txt"))
    assert np.array_equal(k, k2) and np.array_equal(e, e2) and (dmin, itv) == (dmin2, itv2)


def test_pair_file_padding_and_dropping():
    metas = formats.read_pair_file(str(FIX / "pair.txt"), nviews=5)
    assert metas == [(0, [10, 1, 9, 12, 10]), (2, [0, 1, 0, 0, 0])]  # view 1 has no source views: dropped
    assert formats.read_pair_file(str(FIX / "pair.txt"), nviews=2)[0] == (0, [10, 1, 9, 12])


def test_stage_projections_and_depth_values():
    cams = [formats.read_cam_file(str(FIX / n)) for n in ("00000000_cam.txt", "00000001_cam.txt")]
    proj = formats.stage_projections([c[1] for c in cams], [c[0] for c in cams])
    assert set(proj) == {"stage1", "stage2", "stage3", "stage4"}
    for name, s in (("stage1", 0.5), ("stage2", 1.0), ("stage3", 2.0), ("stage4", 4.0)):
        m = proj[name]
        assert m.shape == (2, 2, 4, 4) and m.dtype == np.float32
        assert np.array_equal(m[:, 0], np.stack([c[1] for c in cams]))                       # extrinsics untouched
        assert np.allclose(m[0, 1, :2, :3], cams[0][0][:2] * s, rtol=0, atol=0)               # rows 0-1 scaled (powers of two: exact)
        assert np.array_equal(m[0, 1, 2, :3], cams[0][0][2]) and not m[:, 1, 3].any() and not m[:, 1, :, 3].any()
    dv = formats.depth_values(425.0, 2.5 * 1.06, 192)
    assert dv.dtype == np.float32 and len(dv) == 192 and dv[0] == 425.0
    assert np.array_equal(dv, np.arange(425.0, 2.5 * 1.06 * 191.5 + 425.0, 2.5 * 1.06, dtype=np.float32))


@pytest.mark.skipif(not os.path.isdir("/root/reference/datasets"), reason="reference tree not mounted (GPU box)")
def test_against_the_reference_functions_directly(tmp_path):
    sys.path.insert(0, "/root/reference")
    try:
        from datasets.data_io import read_pfm as ref_read, save_pfm as ref_save
    finally:
        sys.path.pop(0)
    rng = np.random.RandomState(3)
    for shape in ((5, 7), (3, 4, 3), (8, 2, 1)):
        img = rng.randn(*shape).astype(np.float32)
        ref_save(str(tmp_path / "r.pfm"), img, scale=3)
        formats.save_pfm(str(tmp_path / "o.pfm"), img, scale=3)
        assert (tmp_path / "r.pfm").read_bytes() == (tmp_path / "o.pfm").read_bytes()
        a, sa = ref_read(str(tmp_path / "o.pfm"))
        b, sb = formats.read_pfm(str(tmp_path / "r.pfm"))
        assert sa == sb == 3.0 and np.array_equal(np.asarray(a), np.asarray(b))


def _write_scan(root, n_views=4, hw=((140, 200), (140, 200), (150, 210), (140, 200))):
    """A tiny scan directory in the layout the reference's evaluation loader reads (images/, cams/, pair.txt)."""
    from PIL import Image
    from mvster_b200 import synth
    scan = root / "scan1"
    (scan / "images").mkdir(parents=True)
    (scan / "cams").mkdir()
    rng = np.random.RandomState(11)
    cams = synth.arc_cameras(n_views, 140, 200, 3.0)
    for v in range(n_views):
        h, w = hw[v]
        Image.fromarray((rng.rand(h, w, 3) * 255).astype(np.uint8)).save(scan / "images" / f"{v:0>8}.jpg", quality=95)
        k = cams[v, 1, :3, :3].astype(np.float64).copy()
        k[:2] *= 4.0  # cam files hold full-scale intrinsics; the reader divides by 4
        formats.write_cam_file(str(scan / "cams" / f"{v:0>8}_cam.txt"), cams[v, 0], k, 425.0, 2.5)
    (scan / "pair.txt").write_text("4\n0\n3 1 9.0 2 8.0 3 7.0\n1\n2 0 9.0 2 5.0\n2\n1 0 3.0\n3\n0\n")
    return scan


@pytest.mark.skipif(not os.path.isdir("/root/reference/datasets"), reason="reference tree not mounted (GPU box)")
def test_eval_sample_matches_the_reference_dataset(tmp_path):
    """formats.load_eval_sample vs the reference's MVSDataset.__getitem__ (datasets/general_eval4.py) on a synthetic scan:
    images, the four projection stacks and the hypothesis list, bit for bit; one view has a different size and is resized."""
    _write_scan(tmp_path)
    sys.path.insert(0, "/root/reference")
    try:
        from datasets.general_eval4 import MVSDataset
    finally:
        sys.path.pop(0)
    ds = MVSDataset(str(tmp_path), ["scan1"], "test", 4, 1.06, max_h=1200, max_w=1600, fix_res=False)
    metas = formats.read_pair_file(str(tmp_path / "scan1" / "pair.txt"), 4)
    assert [(m[1], m[2]) for m in ds.metas] == metas
    for idx, (ref_view, src_views) in enumerate(metas):
        want = ds[idx]
        got = formats.load_eval_sample(str(tmp_path), "scan1", ref_view, src_views, 4, 1.06, 1200, 1600)
        assert got["filename"] == want["filename"]
        assert np.array_equal(got["depth_values"], want["depth_values"])
        assert len(got["imgs"]) == len(want["imgs"]) == 4
        for a, b in zip(got["imgs"], want["imgs"]):
            assert a.shape == b.shape == (3, 128, 192) and a.dtype == np.float32 and np.array_equal(a, b)
        for s in ("stage1", "stage2", "stage3", "stage4"):
            assert np.array_equal(got["proj_matrices"][s], want["proj_matrices"][s]), s


def test_backprojection_and_ply_round_trip(tmp_path):
    """fusion.backproject_points against the formulas of test_mvs4.py:384-399 written out in numpy, and the PLY writer / reader."""
    from mvster_b200 import fusion
    from oracle import fusion_oracle
    v = fusion_oracle.synthetic_scene(2, 24, 32, seed=2)[1]
    rng = np.random.RandomState(5)
    mask = rng.rand(24, 32) > 0.6
    img = rng.rand(24, 32, 3).astype(np.float32)
    xyz, rgb = fusion.backproject_points(v["depth"], mask, v["K"], v["E"], img)
    x, y = np.meshgrid(np.arange(32), np.arange(24))
    x, y, depth = x[mask], y[mask], v["depth"][mask]
    xyz_ref = np.matmul(np.linalg.inv(v["K"]), np.vstack((x, y, np.ones_like(x))) * depth)
    want = np.matmul(np.linalg.inv(v["E"]), np.vstack((xyz_ref, np.ones_like(x))))[:3].transpose((1, 0))
    assert xyz.dtype == np.float32 and xyz.shape == (int(mask.sum()), 3)
    assert np.abs(xyz - want.astype(np.float32)).max() <= 1e-4 * np.abs(want).max()
    assert np.array_equal(rgb, (img[mask] * 255).astype(np.uint8))
    formats.save_ply(str(tmp_path / "a.ply"), xyz, rgb)
    head = (tmp_path / "a.ply").read_bytes()[:200].decode("ascii", "ignore")
    assert head.startswith("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\n" % len(xyz)) and "property uchar blue" in head
    xyz2, rgb2 = formats.read_ply(str(tmp_path / "a.ply"))
    assert np.array_equal(xyz2, xyz) and np.array_equal(rgb2, rgb)
    formats.save_ply(str(tmp_path / "b.ply"), xyz)
    xyz3, rgb3 = formats.read_ply(str(tmp_path / "b.ply"))
    assert np.array_equal(xyz3, xyz) and rgb3 is None
