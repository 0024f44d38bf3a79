"""Generate tests/golden/et_backward.npz: gradients of the epipolar-Transformer aggregation w.r.t. the feature maps, taken by
torch autograd through the UNMODIFIED reference ``stagenet.forward`` (models/mvs4net_utils.py:1015-1062; imported from
/root/reference, present only in the build container).  The regulariser passed in only captures the cost volume, so the
fixture isolates warp + aggregation.  They pin the oracle's differentiable formulation and, through it, the backward kernel
(csrc/et_fuse_bwd.cu; tests/test_et_backward_golden.py, tests/test_emu_kernels.py, tests/test_gpu_zzz_et_backward.py).

    python oracle/make_et_backward_golden.py

Test infrastructure; nothing at run time on the GPU box reads /root/reference.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("MVSTER_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))

from models import mvs4net_utils as ref_utils  # noqa: E402  (the reference package)
from mvster_b200 import synth  # noqa: E402

CASES = {  # name: (B, Nv, C, G, D, H, W, step_deg, seed)
    "c8": (1, 3, 8, 4, 4, 8, 16, 2.0, 21),
    "c16": (1, 2, 16, 4, 4, 8, 16, 1.0, 22),
    "c32": (2, 3, 32, 8, 8, 4, 8, 5.0, 23),
    "c64": (1, 3, 64, 8, 8, 4, 8, 1.0, 24),
}


def inputs(B, nv, C, D, H, W, step, seed):
    rng = np.random.RandomState(seed)
    feats = [torch.from_numpy(rng.randn(B, C, H, W).astype(np.float32)) for _ in range(nv)]
    cams = synth.stage_projections(synth.arc_cameras(nv, H, W, step), B, num_stage=1)["stage1"]
    inv = torch.linspace(1 / 935.0, 1 / 425.0, D).reshape(1, D, 1, 1)
    hypo = (1.0 / inv).expand(B, D, H, W) * torch.from_numpy(rng.uniform(0.97, 1.03, (B, D, H, W)).astype(np.float32))
    return feats, cams, hypo.contiguous(), rng


def main():
    blob = {}
    for name, (B, nv, C, G, D, H, W, step, seed) in CASES.items():
        feats, cams, hypo, rng = inputs(B, nv, C, D, H, W, step, seed)
        gout = torch.from_numpy(rng.randn(B, G, D, H, W).astype(np.float32))
        leaves = [f.clone().requires_grad_(True) for f in feats]
        captured = {}

        def regnet(cost):
            captured["cost"] = cost
            return cost.sum(1)

        net = ref_utils.stagenet(inverse_depth=True, mono=False, attn_fuse_d=True, attn_temp=2).train()
        net(leaves, cams, hypo, regnet, 0, group_cor=True, group_cor_dim=G, split_itv=0.5)
        grads = torch.autograd.grad(captured["cost"], leaves, gout)
        blob[f"{name}_shape"] = np.array([B, nv, C, G, D, H, W], np.int64)
        blob[f"{name}_cams"] = cams.numpy()
        blob[f"{name}_hypo"] = hypo.numpy()
        blob[f"{name}_gout"] = gout.numpy()
        blob[f"{name}_cost"] = captured["cost"].detach().numpy()
        for v in range(nv):
            blob[f"{name}_feat{v}"] = feats[v].numpy()
            blob[f"{name}_grad{v}"] = grads[v].numpy()
    out = REPO / "tests" / "golden" / "et_backward.npz"
    np.savez_compressed(out, **blob)
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
