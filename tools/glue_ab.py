#!/usr/bin/env python
"""Kernel-level A/B of the CUDA-core kernels and their default-off variants at the cfg2 shapes (5 views, 512x640): CUDA-event
time per launch (L2 flushed before each, median), GB/s on the algorithmic bytes, and the largest deviation from the default
kernel's output.  One process: the library reads its MVSTER_* switches at every launch.

    python tools/glue_ab.py [--reps 20] > gpurun_out/glue_ab.json
"""
import argparse
import json
import os
import statistics
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from mvster_b200 import _lib, capi, fpn_engine  # noqa: E402

CUDA = torch.cuda.is_available()
DEV = torch.device("cuda", 0) if CUDA else torch.device("cpu")   # cpu: dry run of this script on the test emulation (--small)


def timed(fn, flush, reps):
    ts = []
    out = None
    for _ in range(reps + 3):
        if not CUDA:
            t0 = time.perf_counter()
            out = fn()
            ts.append((time.perf_counter() - t0) * 1e6)
            continue
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[3:]), out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--small", action="store_true", help="tiny shapes (dry run)")
    args = ap.parse_args()
    if not CUDA:  # dry run without a GPU: the kernels' sources on the CPU emulation of tests/emu (not a measurement)
        sys.path.insert(0, str(REPO / "tests" / "emu"))
        import install
        install.install()
    g = torch.Generator(device="cpu").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(DEV)
    flush = torch.empty(512 * 1024 * 1024 // 4 if CUDA else 1, dtype=torch.float32, device=DEV)
    lib = _lib.load()
    N, H, W = (1, 32, 64) if args.small else (5, 512, 640)
    cases = []

    U, c0, wc, bt = r(9, N, H // 2, W // 2, 8), r(N, H, W, 8), r(9, 8, 8) / 8, r(9, 8)
    out = torch.empty(N, H, W, 8, device=DEV)

    def gather():
        _lib.check(lib.mvster_fpn_out4_gather_f32(capi._ptr(U), 8, capi._ptr(c0), capi._ptr(wc), capi._ptr(bt), capi._ptr(out), N, H, W,
                                                  capi._stream()), "gather")
        return out
    cases.append((f"fpn_out4_gather ({N} x {H} x {W})", "MVSTER_FPN_GATHER", ["1", "2"], gather, (U.numel() + c0.numel() + out.numel()) * 4))

    for (h, w, CL) in ((H // 2, W // 2, 16), (H // 4, W // 4, 32)):
        top, lat, wl, bl = r(N, h // 2, w // 2, 64), r(N, h, w, CL), r(CL, 64) / 4, r(64)
        cases.append((f"fpn_merge<{CL}> ({N} x {h} x {w})", "MVSTER_FPN_MERGE", ["1", "2", "3"],
                      (lambda top=top, lat=lat, wl=wl, bl=bl: fpn_engine._merge(top, lat, wl, bl)),
                      (top.numel() + lat.numel() + N * h * w * 64) * 4))

    img, ws, bs = torch.rand(N, 3, H, W, generator=g).to(DEV), r(9, 3, 8) / 3, r(8)
    o8 = torch.empty(N, H, W, 8, device=DEV)

    def stem():
        _lib.check(lib.mvster_conv_first_f32(capi._ptr(img), capi._ptr(ws), capi._ptr(bs), capi._ptr(o8), N, H, W, capi._stream()), "stem")
        return o8
    cases.append((f"conv_first ({N} x {H} x {W})", "MVSTER_CONV_FIRST", ["1", "2"], stem, (img.numel() + o8.numel()) * 4))

    for (D, h, w, G) in ((4, H, W, 4), (4, H // 2, W // 2, 4), (8, H // 4, W // 4, 8), (8, H // 8, W // 8, 8)):
        x, w0, b0 = r(1, D, h, w, G), r(9, G, 8) / 3, r(8)
        cases.append((f"reg2d conv0 G={G} ({D} x {h} x {w})", "MVSTER_CONV0_PX4", ["0", "1"],
                      (lambda x=x, w0=w0, b0=b0: capi.conv3d_ndhwc(x, w0, b0, 1)), (x.numel() + D * h * w * 8) * 4))

    rows = []
    for name, key, values, fn, nbytes in cases:
        base = None
        for v in values:
            os.environ.pop(key, None)
            if v is not None:
                os.environ[key] = v
            us, o = timed(fn, flush, args.reps)
            o = o.clone()
            if base is None:
                base = o
            dev = ((o - base).abs().max() / base.abs().max()).item()
            rows.append({"kernel": name, "switch": f"{key}={v}" if v else "default", "us": round(us, 2), "GB/s": round(nbytes / us / 1e3, 1),
                         "max_dev_vs_default": dev})
        os.environ.pop(key, None)
    print(json.dumps({"workload": "cfg2 shapes, fp32, L2 flushed before every launch" if CUDA and not args.small else "dry run", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
