#!/usr/bin/env python
"""Time of the geometric-consistency kernel (csrc/fusion.cu) per (reference, source) pair against its HBM
roofline: per reference pixel it reads the reference depth (4 B) and a 2 x 2 patch of the source depth (hits in L1 / L2:
the source map is read about once, 4 B), and writes the reprojected depth, the sampling coordinates, the mask (4 + 4 + 4 + 1 B)
and - in the accumulating form used by fuse_reference_view - updates the view count and the depth sum (2 x (4 + 4) B): 21 B per
pixel for check_geometric_consistency, 24 B for the accumulating launch.

    python tools/fusion_bench.py > gpurun_out/fusion_bench.json
"""
import json
import statistics
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from mvster_b200 import fusion, synth  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peak = 6549.1
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    rows = []
    for H, W in ((512, 640), (1152, 1600)):
        cams = synth.arc_cameras(2, H, W, 1.0)
        rng = np.random.RandomState(0)
        d_ref = torch.from_numpy(rng.uniform(500, 900, (H, W)).astype(np.float32)).to(dev)
        d_src = torch.from_numpy(rng.uniform(500, 900, (H, W)).astype(np.float32)).to(dev)
        mats = fusion._prepare_mats(cams[0, 1], cams[0, 0], cams[1, 1], cams[1, 0])
        count = torch.zeros((H, W), device=dev, dtype=torch.int32)
        dsum = torch.zeros((H, W), device=dev, dtype=torch.float32)
        for name, fn, bpp in (("check_geometric_consistency (coordinates + mask out)", lambda: fusion._launch(d_ref, d_src, mats, True, True), 21),
                              ("accumulating launch of fuse_reference_view", lambda: fusion._launch(d_ref, d_src, mats, False, False, count, dsum), 24)):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            us = statistics.median(ts)
            gbs = H * W * bpp / (us * 1e-6) / 1e9
            rows.append({"H": H, "W": W, "launch": name, "us_incl_output_allocation": round(us, 2), "bytes_per_pixel": bpp,
                         "GB/s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3)})
    print(json.dumps({"peak_gbs": peak, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
