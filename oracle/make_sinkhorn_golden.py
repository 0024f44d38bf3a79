"""Generates tests/golden/sinkhorn.npz by running the UNMODIFIED reference `sinkhorn` (models/mvs4net_utils.py:1096-1142), loss and
autograd gradient w.r.t. the predicted distribution, on seeded synthetic inputs.  Run in the build container (needs
/root/reference); the fixtures travel, the reference does not.

    python oracle/make_sinkhorn_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")
from models.mvs4net_utils import sinkhorn  # noqa: E402

CASES = [  # name, B, D, H, W, iters, eps, continuous
    ("d8_it3", 2, 8, 5, 7, 3, 1.0, False),
    ("d8_it10", 1, 8, 6, 9, 10, 1.0, False),
    ("d4_it10", 2, 4, 4, 10, 10, 1.0, False),
    ("d4_it10_eps", 1, 4, 6, 6, 10, 0.7, False),
    ("d8_it10_cont", 1, 8, 5, 8, 10, 1.0, True),
    ("d4_it5_cont", 2, 4, 4, 9, 5, 1.0, True),
]


def main():
    out = {}
    for idx, (name, B, D, H, W, iters, eps, cont) in enumerate(CASES):
        g = torch.Generator().manual_seed(1000 + idx)
        inv = torch.linspace(1 / 935.0, 1 / 425.0, D).reshape(1, D, 1, 1) * (1 + 0.02 * torch.rand(B, 1, H, W, generator=g))
        hypo = (1.0 / inv).contiguous()                                   # far -> near, like the inverse-depth schedule
        gt = (1.0 / (inv[:, 0] + (inv[:, -1] - inv[:, 0]) * torch.rand(B, H, W, generator=g) * 1.1)).contiguous()
        attn = torch.softmax(torch.randn(B, D, H, W, generator=g) * 2, 1).requires_grad_(True)
        mask = torch.rand(B, H, W, generator=g) > 0.25
        _, loss = sinkhorn(gt, hypo, attn, mask, iters, eps=eps, continuous=cont)
        (grad,) = torch.autograd.grad(loss, attn)
        for k, v in dict(gt=gt, hypo=hypo, attn=attn.detach(), mask=mask, loss=loss.detach(), grad=grad).items():
            out[f"{name}_{k}"] = v.numpy()
        out[f"{name}_meta"] = np.array([B, D, H, W, iters, eps, int(cont)], dtype=np.float64)
    np.savez_compressed(REPO / "tests" / "golden" / "sinkhorn.npz", **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
