#!/usr/bin/env python
"""Forward + backward of the warp + ET aggregation at the cfg2 stage shapes (5 views, 512x640): the autograd node over the
fused kernels (mvster_b200/train_ops.py) against the PyTorch-op formulation the reference trains with
(mvster_b200/torch_path.aggregate).  Per stage: CUDA-event time of forward and of backward (L2 flushed before each), peak
extra device memory between forward and backward, and the largest gradient deviation between the two.

    python tools/et_bwd_bench.py [--reps 10] > gpurun_out/et_bwd_bench.json
"""
import argparse
import json
import statistics
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from mvster_b200 import synth, torch_path, train_ops  # noqa: E402

STAGES = [(64, 8, 8, 3), (32, 8, 8, 2), (16, 4, 4, 1), (8, 4, 4, 0)]  # (C, G, D, log2 down-scale)


def timed(fn, flush, reps):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B, NV, H, W = 1, 5, 512, 640
    _, proj, _ = synth.make_inputs(B, NV, H, W, seed=0)
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    rows = []
    for k, (C, G, D, s) in enumerate(STAGES):
        h, w = H >> s, W >> s
        g = torch.Generator(device="cpu").manual_seed(k)
        feats = [torch.randn(B, C, h, w, generator=g).to(dev) for _ in range(NV)]
        cams = proj[f"stage{k + 1}"].to(dev)
        centre = 450.0 + 450.0 * torch.rand(B, 1, h, w, generator=g)
        span = [0.5, 0.12, 0.03, 0.015][k]
        hypo = (centre * (1.0 + span * torch.linspace(0.5, -0.5, D).reshape(1, D, 1, 1))).to(dev).contiguous()
        gout = torch.randn(B, G, D, h, w, generator=g).to(dev)
        row = {"stage": k + 1, "C": C, "G": G, "D": D, "H": h, "W": w}
        grads = {}
        for name, fn in (("kernels", lambda f: train_ops.aggregate(f, cams, hypo, G, 2.0)),
                         ("pytorch", lambda f: torch_path.aggregate(f, cams, hypo, True, G, 2.0, True))):
            leaves = [f.clone().requires_grad_(True) for f in feats]
            for _ in range(2):  # warm-up
                torch.autograd.grad(fn(leaves), leaves, gout)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
            t_f, cost = timed(lambda: fn(leaves), flush, args.reps)
            held = torch.cuda.memory_allocated() - base          # what autograd keeps alive for the backward (one graph)
            t_b, gr = timed(lambda: torch.autograd.grad(cost, leaves, gout, retain_graph=True), flush, args.reps)
            grads[name] = [x.detach() for x in gr]
            row[name] = {"fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4), "held_MB": round(held / 1e6, 1),
                         "peak_MB": round((torch.cuda.max_memory_allocated() - base) / 1e6, 1)}
            del cost, gr, leaves
        row["grad_max_rel_dev"] = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(grads["kernels"], grads["pytorch"]))
        rows.append(row)
    print(json.dumps({"workload": "cfg2 stage shapes, 5 views 512x640, fp32", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
