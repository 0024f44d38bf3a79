#!/usr/bin/env python
"""A/B timing of the warp + ET kernel variants on the cfg2 stage inputs (one process: the library reads its
MVSTER_ET_* switches at every launch).  For each variant: CUDA-event time per stage (L2 flushed before every launch),
GB/s on the algorithmic bytes, and the largest deviation from the per-hypothesis tiled kernel's cost volume.

    python tools/et_ab.py [--reps 20] > gpurun_out/et_ab.json
"""
import argparse
import json
import os
import statistics
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
import bench  # noqa: E402
from mvster_b200 import capi, synth  # noqa: E402

VARIANTS = [
    ("tiled", dict(MVSTER_ET_WIN="0")),
    ("tma_15x32", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="1", MVSTER_ET_TMA_TH="15")),
    ("tma_7x32", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="1", MVSTER_ET_TMA_TH="7")),
    ("tma_15x32_interleaved*", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="1", MVSTER_ET_TMA_TH="15", IL="1")),
    ("tma_7x32_interleaved*", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="1", MVSTER_ET_TMA_TH="7", IL="1")),
    ("win", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="0")),
    ("win_interleaved*", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="0", IL="1")),
    ("win_mb5", dict(MVSTER_ET_WIN="1", MVSTER_ET_TMA="0", MVSTER_ET_WIN_MB="5")),
]
KEYS = ("MVSTER_ET_WIN", "MVSTER_ET_WIN_MB", "MVSTER_ET_PREFETCH", "MVSTER_ET_TMA", "MVSTER_ET_TMA_TH", "MVSTER_ET_TMA_DEBUG", "MVSTER_ET_TMA_BOXES")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, NV, H, W = 1, 5, 512, 640
    model = bench.build_model(dev)
    imgs, proj, dv = synth.make_inputs(B, NV, H, W, seed=0)
    imgs_d = [t.to(dev) for t in imgs]
    proj_d = {k: v.to(dev) for k, v in proj.items()}
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    with torch.no_grad():
        out = model(imgs_d, proj_d, dv.to(dev))
        x = torch.cat(imgs_d, 0).contiguous(memory_format=torch.channels_last)
        pyr = model.feature(x)
    stages = []
    for k in range(4):
        f = capi.to_nhwc(pyr[f"stage{k + 1}"])
        feats = [f[v * B:(v + 1) * B] for v in range(NV)]
        stages.append((feats, out[f"stage{k + 1}"]["hypo_depth"], capi.pose(proj_d[f"stage{k + 1}"])))
    res, base = {}, [None] * 4
    for name, env in VARIANTS:
        for key in KEYS:
            os.environ.pop(key, None)
        env = dict(env)
        il = env.pop("IL", "0") == "1"  # * = timing only: the features here are in natural order, so the values are not comparable
        os.environ.update(env)
        row = []
        for k, (feats, hypo, pose) in enumerate(stages):
            G = bench.G_K[k]
            kw = dict(interleaved=il and k > 0)
            cost = torch.empty((B, bench.D_K[k], H >> (3 - k), W >> (3 - k), G), device=dev)
            for _ in range(3):
                capi.et_fuse(feats[0], feats[1:], pose, hypo, G, 2.0, cost=cost, **kw)
            ts = []
            for _ in range(args.reps):
                flush.fill_(1.0)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                capi.et_fuse(feats[0], feats[1:], pose, hypo, G, 2.0, cost=cost, **kw)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            if base[k] is None:
                base[k] = cost.clone()
            dev_max = (cost - base[k]).abs().max().item() / base[k].abs().max().item()
            nbytes = bench.et_algorithmic_bytes(k, B, NV - 1, H, W)
            t = statistics.mean(ts)
            row.append({"stage": k + 1, "us": round(t * 1e3, 2), "min_us": round(min(ts) * 1e3, 2),
                        "gbs": round(nbytes / (t * 1e-3) / 1e9, 1), "rel_dev_vs_tiled": dev_max})
        res[name] = row
        print(name, capi.et_last_kernel(), [r["us"] for r in row], "dev", ["%.1e" % r["rel_dev_vs_tiled"] for r in row], file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
