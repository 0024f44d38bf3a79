"""Golden vectors for the geometric-consistency filter, produced by the UNMODIFIED reference: the function sources of
reproject_with_depth / check_geometric_consistency are cut out of /root/reference/test_mvs4.py (the module itself cannot be
imported here: plyfile / tensorboardX are absent) and executed as they are.  Run in the build container:
    python oracle/make_fusion_golden.py
"""
import ast
import sys
from pathlib import Path

import cv2
import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import fusion_oracle  # noqa: E402

src = Path("/root/reference/test_mvs4.py").read_text()
tree = ast.parse(src)
ns = {"np": np, "cv2": cv2}
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("reproject_with_depth", "check_geometric_consistency"):
        exec(compile(ast.Module([node], []), "test_mvs4.py", "exec"), ns)

out = REPO / "tests" / "golden" / "fusion"
out.mkdir(parents=True, exist_ok=True)
for name, (nv, H, W, seed, step) in {"plane_4v_48x64": (4, 48, 64, 0, 4.0), "plane_3v_40x56_wide": (3, 40, 56, 1, 9.0)}.items():
    views = fusion_oracle.synthetic_scene(nv, H, W, seed, step)
    ref = views[0]
    rec = {"meta": np.array([nv, H, W, seed], np.int64), "step": np.array(step)}
    masks, depths = [], []
    for v in range(1, nv):
        s = views[v]
        with np.errstate(divide="ignore", invalid="ignore"):
            mask, dr, xs, ys = ns["check_geometric_consistency"](ref["depth"].copy(), ref["K"], ref["E"], s["depth"].copy(), s["K"], s["E"])
        rec[f"mask{v}"], rec[f"depth_reprojected{v}"], rec[f"x2d_src{v}"], rec[f"y2d_src{v}"] = mask, dr, xs, ys
        masks.append(mask.astype(np.int32))
        depths.append(dr)
    # the averaging of filter_depth (test_mvs4.py:375-378), thres_view = 2, conf = 0.5 on a synthetic confidence map
    geo_sum = sum(masks)
    rec["geo_mask_sum"] = geo_sum
    rec["depth_est_averaged"] = (sum(depths) + ref["depth"]) / (geo_sum + 1)
    conf = np.random.RandomState(seed + 10).rand(H, W).astype(np.float32)
    rec["confidence"] = conf
    rec["final_mask"] = np.logical_and(conf > 0.5, geo_sum >= 2)
    np.savez_compressed(out / f"{name}.npz", **rec)
    print(name, "mask fractions", [float(m.mean()) for m in masks], "final", float(rec["final_mask"].mean()))
