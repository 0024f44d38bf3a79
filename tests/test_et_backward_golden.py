"""Gradients of the warp + epipolar-Transformer aggregation w.r.t. the feature maps: the oracle's and the package's
differentiable PyTorch formulations against gradients taken through the unmodified reference's ``stagenet.forward``
(tests/golden/et_backward.npz, written by oracle/make_et_backward_golden.py).  The backward KERNEL is checked against the
same vectors in tests/test_emu_kernels.py (CPU emulation) and tests/test_gpu_zzz_et_backward.py (B200)."""
import numpy as np
import pytest
import torch

from util import GOLDEN, oracle

from mvster_b200 import torch_path

CASES = ["c8", "c16", "c32", "c64"]


def load_case(name):
    z = np.load(GOLDEN / "et_backward.npz")
    B, nv, C, G, D, H, W = (int(v) for v in z[f"{name}_shape"])
    t = lambda k: torch.from_numpy(z[f"{name}_{k}"])
    return dict(G=G, nv=nv, feats=[t(f"feat{v}") for v in range(nv)], cams=t("cams"), hypo=t("hypo"), gout=t("gout"),
                cost=t("cost"), grads=[t(f"grad{v}") for v in range(nv)])


@pytest.mark.parametrize("name", CASES)
def test_oracle_and_torch_path_gradients_match_reference(name):
    c = load_case(name)
    for fn in (lambda f: oracle.et_aggregate(f, c["cams"], c["hypo"], True, c["G"], 2.0),
               lambda f: torch_path.aggregate(f, c["cams"], c["hypo"], True, c["G"], 2.0, True)):
        leaves = [f.clone().requires_grad_(True) for f in c["feats"]]
        cost = fn(leaves)
        assert (cost.detach() - c["cost"]).abs().max().item() <= 2e-5 * c["cost"].abs().max().item()
        grads = torch.autograd.grad(cost, leaves, c["gout"])
        for v in range(c["nv"]):
            scale = c["grads"][v].abs().max().item()
            assert (grads[v] - c["grads"][v]).abs().max().item() <= 2e-5 * scale
