"""Forward + backward of the optimal-transport loss term at the cfg2 stage shapes: fused kernel (csrc/sinkhorn.cu) against the PyTorch ops
of mvster_b200/losses.py (CUDA events, 5 repetitions after 2 warm-ups).  python tools/sinkhorn_bench.py [iters]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mvster_b200 import losses, train_ops  # noqa: E402


def timed(fn, n=5):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    out = {"iters": iters, "stages": []}
    for D, h, w in ((8, 64, 80), (8, 128, 160), (4, 256, 320), (4, 512, 640)):
        inv = torch.linspace(1 / 935.0, 1 / 425.0, D, device="cuda").reshape(1, D, 1, 1) * (1 + 0.02 * torch.rand(1, 1, h, w, device="cuda"))
        hypo = 1.0 / inv
        gt = 1.0 / (inv[:, 0] + (inv[:, -1] - inv[:, 0]) * torch.rand(1, h, w, device="cuda"))
        mask = torch.rand(1, h, w, device="cuda") > 0.3
        attn = torch.softmax(torch.randn(1, D, h, w, device="cuda") * 2, 1).requires_grad_(True)

        def fused():
            torch.autograd.grad(train_ops.sinkhorn_loss(gt, hypo, attn, mask, iters, 1.0, False), attn)

        def ops():
            torch.autograd.grad(losses.sinkhorn(gt, hypo, attn, mask, iters, 1.0, False)[1], attn)
        out["stages"].append({"D": D, "H": h, "W": w, "fused_ms": round(timed(fused), 4), "torch_ops_ms": round(timed(ops), 4)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
