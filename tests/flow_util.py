"""The evaluation flow of test_mvs4.py:199-262 (save_depth) and :332-421 (filter_depth) on a synthetic scan in the reference's
directory layout, built from this package's pieces: prefetch -> forward -> PFM depth / confidence -> geometric-consistency
filter -> back-projection -> PLY.  Shared by the GPU test and the CPU-emulation test; ``forward(sample) -> outputs``."""
import numpy as np

from test_formats import _write_scan

from mvster_b200 import formats, fusion, prefetch


def scan_to_point_cloud(tmp_path, forward, device, hw=(140, 200)):
    _write_scan(tmp_path, hw=(hw,) * 4)
    size = (hw[0] // 64 * 64, hw[1] // 64 * 64)               # the loader shrinks both sides to multiples of 64
    pairs = formats.read_pair_file(str(tmp_path / "scan1" / "pair.txt"), 4)
    out_dir = tmp_path / "out"
    depths, confs, cams = {}, {}, {}
    for (ref, _), s in zip(pairs, prefetch.Prefetcher(prefetch.eval_jobs(str(tmp_path), "scan1", pairs, 4), device=device)):
        out = forward(s)
        d, c = out["depth"][0].cpu().numpy(), out["photometric_confidence"][0].cpu().numpy()
        assert d.shape == c.shape == size and np.isfinite(d).all() and (d > 0).all()
        assert 0.0 <= c.min() and c.max() <= 1.0 + 1e-6
        name = s["filename"][0]
        for kind, arr in (("depth_est", d), ("confidence", c)):
            path = out_dir / name.format(kind, ".pfm")
            path.parent.mkdir(parents=True, exist_ok=True)
            formats.save_pfm(str(path), arr)
        back, _ = formats.read_pfm(str(out_dir / name.format("depth_est", ".pfm")))
        assert np.array_equal(back, d)
        P = s["proj_matrices"]["stage4"][0].cpu().numpy()
        depths[ref], confs[ref], cams[ref] = d, c, (P[0, 1, :3, :3].copy(), P[0, 0].copy())
    assert sorted(depths) == [0, 1, 2]
    ref = pairs[0][0]
    srcs = [v for v in dict.fromkeys(pairs[0][1]) if v in depths]
    assert srcs == [1, 2]
    views = [(depths[v], cams[v][0], cams[v][1]) for v in srcs]
    res = fusion.fuse_reference_view(depths[ref], cams[ref][0], cams[ref][1], confs[ref], views, conf_thres=0.0, thres_view=1,
                                     device=device)
    assert res["depth_est_averaged"].shape == size and res["final_mask"].dtype == bool
    assert res["geo_mask_sum"].max() <= len(srcs)
    xyz, _ = fusion.backproject_points(res["depth_est_averaged"], res["final_mask"], cams[ref][0], cams[ref][1])
    assert xyz.shape == (int(res["final_mask"].sum()), 3) and np.isfinite(xyz).all()
    formats.save_ply(str(out_dir / "scan1.ply"), xyz)
    got, _ = formats.read_ply(str(out_dir / "scan1.ply"))
    assert np.array_equal(got, xyz.astype(np.float32))
    return res
